"""Drop-in for the reference's skeleton block_6_admm_loop.py (signature :72-84, history aliases :101-105):
same entry point as block_6_admm_loop_ver2; the scs_* kwargs are accepted and ignored."""
from block_6_admm_loop_ver2 import decentralized_admm  # noqa: F401
