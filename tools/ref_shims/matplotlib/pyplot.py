"""Empty stand-in (see __init__)."""
