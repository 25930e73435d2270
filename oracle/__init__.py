"""CPU oracle for soundscope's analyzer hot path — TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (soundscope_b200) never does.
"""
from .binding import *  # noqa: F401,F403
from . import capture_ref  # noqa: F401,E402  (PCM conversion, capture ring, microphone tick: numpy)
