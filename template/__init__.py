"""Legacy version module read by the launch scripts' auto-updater (reference template/__init__.py:24-27)."""
__version__ = "0.1.0"
version_split = __version__.split(".")
__spec_version__ = (1000 * int(version_split[0])) + (10 * int(version_split[1])) + (1 * int(version_split[2]))
