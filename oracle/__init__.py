"""CPU oracle for the slab-decomposed 3-D C2C FFT hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` arm may
import this package.  The product (distributedfft_b200) never does.
"""
from .oracle import *  # noqa: F401,F403
