"""Empty stand-in: only the reference visualisation helpers touch imageio."""
