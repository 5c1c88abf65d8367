"""Multi-process plumbing: one process per GPU, QPs connected across ranks over NVLink (CUDA IPC)."""
from .peer import PeerInfo, RemoteMR, connect_ring, describe_local, connect_to  # noqa: F401
