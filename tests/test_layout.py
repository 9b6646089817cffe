"""Repository rules: the product never imports the oracle; required files exist."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sde4mbrl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src.replace("oracle/__init__.py", ""), f


def test_required_files():
    for f in ("bench.py", "__graft_entry__.py", "DESIGN.md", "INTEGRATION.md", "include/sde4b200.h",
              "oracle/__init__.py", "tests/golden/make_golden.py"):
        assert os.path.exists(os.path.join(ROOT, f)), f


def test_oracle_headers_say_test_infrastructure():
    for f in ("oracle/__init__.py", "oracle/threefry.py", "oracle/nsde_oracle.py", "oracle/apg_oracle.py",
              "oracle/c/sde_oracle_core.h"):
        head = open(os.path.join(ROOT, f)).read(2500).lower()
        assert "test infrastructure" in head, f
