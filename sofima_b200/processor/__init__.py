"""SubvolumeProcessor plugins on the B200 backend (reference `processor/`)."""
