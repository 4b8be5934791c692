"""Volume decorators on the B200 backend (reference `decorators/flow.py`)."""
