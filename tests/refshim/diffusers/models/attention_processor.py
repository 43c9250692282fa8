class Attention:  # only used as a type annotation by the reference (p2p/ptp_utils.py:29,40)
    pass
