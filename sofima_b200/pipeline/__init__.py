"""Pipeline configurations (reference pipeline/)."""
