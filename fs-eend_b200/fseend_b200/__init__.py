"""fseend_b200: ctypes binding (native.py) and in-tree build (build.py) of the sm_100a FS-EEND hot path."""
from .build import LIB_PATH, build  # noqa: F401
