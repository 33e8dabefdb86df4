"""matplotlib stand-in (imported, unused: motionclone_functions.py:5-6)."""
