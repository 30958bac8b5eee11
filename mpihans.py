#!/usr/bin/env python
"""`python mpihans.py < input.txt > log.out` -- same command line as the reference's mpihans.py
(/root/reference/README.md:12); the implementation is hans_b200/mpihans.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hans_b200.mpihans import cli  # noqa: E402

if __name__ == "__main__":
    cli()
