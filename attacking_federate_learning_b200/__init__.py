"""B200-native Byzantine-robust aggregation engine (drop-in behind the reference's defences.py /
malicious.py / Server.defend).  See DESIGN.md and include/afl_b200.h."""
from . import defences, malicious            # noqa: F401
from .defences import DefenseTypes, defend   # noqa: F401

__all__ = ["defences", "malicious", "DefenseTypes", "defend"]
