"""Package version (reference ``torchrec/version.py``)."""
__version__ = "0.1.0"
github_version = "0.1.0"
