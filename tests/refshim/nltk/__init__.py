"""Import shim: the reference imports nltk at module scope in p2p/ptp_controller_utils.py:6-7; only the demo
driver's heuristic (`preprocessing`) calls it."""


def download(*a, **k):
    return True
