#!/usr/bin/env python3
"""Writes the k-vector files the reference's InitSystem() reads into <out> (oracle/_ref/config): the generator itself
lives in tdvmc_b200/driver.py (write_kvectors) because the GPU-bound driver needs the same files."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tdvmc_b200.driver import write_kvectors  # noqa: E402

write_kvectors(sys.argv[1])
