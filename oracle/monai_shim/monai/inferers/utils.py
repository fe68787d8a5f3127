from oracle.monai08 import _get_scan_interval  # noqa: F401
