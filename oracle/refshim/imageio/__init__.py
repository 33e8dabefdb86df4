"""imageio stand-in (mp4 writing is off the path)."""
