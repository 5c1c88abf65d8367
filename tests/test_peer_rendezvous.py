"""Out-of-band QP exchange (parallel/peer.py) over a real process group: gloo, 3 processes, no GPU.
The CUDA-IPC parts are replaced by stand-ins; what is tested is what travels between ranks and who connects
to whom (each rank must receive exactly its ring successor's description, byte-exact handles included)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

WORKER = textwrap.dedent('''
    import os, sys, json
    import torch.distributed as dist
    sys.path.insert(0, os.environ["RN_REPO"])
    from rocnrdma_b200.parallel import peer

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}", rank=rank, world_size=world)

    def describe(ctx, qp, mrs, r):
        info = peer.PeerInfo(rank=r, device=r, arena_handle=bytes([r]) * 64, arena_base=0x7000_0000_0000 + (r << 32), arena_size=1 << 24,
                             qpn=100 + r, rq=0x7000_0000_1000, rq_dbr=0x7000_0000_2000, rq_log=8, rcq=0x7000_0000_3000,
                             rcq_buf=0x7000_0000_4000, n_mkeys=64)
        for k, (addr, length) in enumerate(mrs):
            info.mrs.append(dict(key=((k + 1) << 8) | r, addr=addr, length=length, access=0xf, handle=bytes([0x80 | r]) * 64,
                                 alloc_base=addr & ~0xfffff, alloc_size=1 << 21))
        return info

    seen = {}
    def connect(ctx, qp, info):
        seen["info"] = info
        return {m["key"]: peer.RemoteMR(addr=m["addr"], length=m["length"], rkey=m["key"]) for m in info.mrs}

    mrs = [(0x7f00_0000_0000 + (rank << 30), 1 << 20), (0x7f10_0000_0000 + (rank << 30), 4096)]
    nxt, remote = peer.connect_ring(None, None, mrs, describe=describe, connect=connect)
    info = seen["info"]
    out = dict(rank=rank, nxt=nxt, peer_rank=info.rank, qpn=info.qpn, arena_handle_ok=info.arena_handle == bytes([nxt]) * 64,
               mr_handle_ok=all(m["handle"] == bytes([0x80 | nxt]) * 64 for m in info.mrs),
               rkeys=sorted(remote), addrs=[remote[k].addr for k in sorted(remote)], key_is_rkey=all(v.key == v.rkey == k for k, v in remote.items()))
    print("RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()
''')


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_ring_rendezvous_over_gloo_three_processes(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    world, port = 3, _free_port()
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RN_REPO=repo,
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    results = {}
    for r, p in enumerate(procs):
        try:
            out, err = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("rendezvous timed out")
        assert p.returncode == 0, err[-2000:]
        line = [l for l in out.splitlines() if l.startswith("RESULT ")][-1]
        results[r] = json.loads(line[len("RESULT "):])
    for r in range(world):
        res, nxt = results[r], (r + 1) % world
        assert res["nxt"] == nxt and res["peer_rank"] == nxt and res["qpn"] == 100 + nxt
        assert res["arena_handle_ok"] and res["mr_handle_ok"] and res["key_is_rkey"]
        assert res["rkeys"] == [(1 << 8) | nxt, (2 << 8) | nxt]
        assert res["addrs"] == [0x7f00_0000_0000 + (nxt << 30), 0x7f10_0000_0000 + (nxt << 30)]
