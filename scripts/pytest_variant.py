#!/usr/bin/env python
"""Developer experiment: run pytest against a variant build (openkite_b200/_variants/<tag>, see build.py --variant).
Usage: python scripts/pytest_variant.py <tag> <pytest args...>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openkite_b200.engine as e
e.LIB_PATH = os.path.join(ROOT, "openkite_b200", "_variants", sys.argv[1], "libkite_b200.so")
import pytest
sys.exit(pytest.main(sys.argv[2:]))
