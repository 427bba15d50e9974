"""Builds oracle/libtiporacle.so (gcc, -ffp-contract=off).  Test infrastructure only."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import c_oracle  # noqa: E402

if __name__ == "__main__":
    print(c_oracle.build(force="--force" in sys.argv))
