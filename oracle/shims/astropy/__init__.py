"""Test-infrastructure shim: lets the unmodified reference import `astropy.constants.G`
in an image without astropy.  Not product code; only oracle/ and tests/ put this on sys.path."""
