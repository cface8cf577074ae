"""Drop-in for the reference's `ocsort_tracker` package (`from ocsort_tracker import ocsort`, clearcam.py:10)."""
from . import ocsort  # noqa: F401
from .STrack import STrack  # noqa: F401
