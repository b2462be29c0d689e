"""ORACLE — test infrastructure only (see oracle/unet.py).  Never imported by `storygen_b200`."""
