"""Loads the UNMODIFIED reference sampler modules from /root/reference/text-guided (this container only; the GPU
box has no /root/reference) behind the import shims in tests/refshim.  Used by tests/test_oracle_pin.py and
tests/make_golden.py to pin the oracle."""
import os
import sys

REF_ROOT = "/root/reference/text-guided"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def reference_available() -> bool:
    return os.path.isdir(REF_ROOT)


def load_reference():
    """Returns a namespace with the reference callables on the north-star path."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    for p in (_SHIM, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib

    ns = type("Ref", (), {})()
    ns.p2p_h_edit = importlib.import_module("inversion.p2p_h_edit")
    ns.inversion_utils = importlib.import_module("inversion.inversion_utils")
    ns.ddpm_inversion = importlib.import_module("inversion.ddpm_inversion")
    ns.ptp_utils = importlib.import_module("p2p.ptp_utils")
    ns.ptp_classes = importlib.import_module("p2p.ptp_classes")
    ns.ptp_controller_utils = importlib.import_module("p2p.ptp_controller_utils")
    ns.seq_aligner = importlib.import_module("p2p.seq_aligner")
    return ns
