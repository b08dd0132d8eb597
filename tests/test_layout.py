"""Repository rules: the product never touches the oracle, the C ABI is complete."""
import os
import re

from conftest import ROOT


def _py_files(d):
    for base, _, files in os.walk(os.path.join(ROOT, d)):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(base, f)


def test_product_does_not_import_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for path in _py_files("aps_b200"):
        assert not pat.search(open(path).read()), f"{path} imports the oracle"


def test_product_never_reads_reference_tree():
    # (__graft_entry__.build() may look for the tree to build oracle/_ref; smoke() and bench.py may not)
    for path in list(_py_files("aps_b200")) + [os.path.join(ROOT, "bench.py")]:
        if os.path.exists(path):
            assert "/root/reference" not in open(path).read().replace("/root/reference/aps", "").replace(
                "(/root/reference", ""), f"{path} touches the reference tree"


def test_header_symbols_exported_and_bound():
    """Every function declared in include/aps_b200.h is exported by the built library and bound in _lib.py."""
    import ctypes

    from aps_b200 import _lib
    header = open(os.path.join(ROOT, "include", "aps_b200.h")).read()
    declared = set(re.findall(r"\b(aps_b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    bound = set(_lib.exported_symbols())
    assert declared == bound, f"header vs binding mismatch: {declared ^ bound}"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert _lib.load().aps_b200_abi_version() == _lib.ABI_VERSION
