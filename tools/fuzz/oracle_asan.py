#!/usr/bin/env python3
"""tools/fuzz/oracle_asan.py <asan build of libpf_oracle.so> — runs the oracle's own tests (fuzzed scenes, strips,
threads, clip cases, dilation, text chain) against a sanitizer build of the oracle. Start it through run.sh, which
builds the library and preloads the sanitizer runtimes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pf_oracle as O  # noqa: E402

O._LIB_PATH = sys.argv[1]
O.build = lambda force=False: O._LIB_PATH
import pytest  # noqa: E402

tests = [os.path.join(ROOT, "tests", t) for t in ("test_oracle.py", "test_dilate_host.py", "test_text_filter_oracle.py")]
sys.exit(pytest.main(tests + ["-q", "-x", "-p", "no:cacheprovider", "--no-header"]))
