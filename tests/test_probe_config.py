"""N0 probe and configuration surface (CPU only; a fake sysfs tree stands in for the GPU box)."""
import os

import pytest

from rocnrdma_b200 import config as Cfg


def test_parse_size_and_sweep():
    assert Cfg.parse_size("4k") == 4096 and Cfg.parse_size("1MiB") == 1 << 20 and Cfg.parse_size("2g") == 2 << 30
    assert Cfg.parse_size("256 MiB") == 256 << 20 and Cfg.parse_size(17) == 17
    with pytest.raises(ValueError):
        Cfg.parse_size("1.5")
    assert Cfg.parse_sweep("1k:8k") == [1024, 2048, 4096, 8192]
    assert Cfg.parse_sweep("1k:1m:x32") == [1024, 32768, 1 << 20]
    assert Cfg.parse_sweep("64,4k") == [64, 4096]
    assert len(Cfg.parse_sweep("1k:1g")) == 21             # the BASELINE sweep: 1 KB - 1 GB


def test_config_from_env_and_validation():
    c = Cfg.Config.from_env({"ROCNRDMA_WIRE": "softhca", "ROCNRDMA_QP_DEPTH": "64", "ROCNRDMA_CHUNK_BYTES": "1m",
                             "ROCNRDMA_AFFINITY": "0:ibp2,1:ibp4"})
    assert c.wire == "softhca" and c.qp_depth == 64 and c.chunk_bytes == 1 << 20 and c.affinity == {0: "ibp2", 1: "ibp4"}
    with pytest.raises(ValueError):
        Cfg.Config.from_env({"ROCNRDMA_QP_DEPTH": "100"})
    with pytest.raises(ValueError):
        Cfg.Config.from_env({"ROCNRDMA_WIRE": "tcp"})


def _fake_sysfs(root):
    def w(path, text):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(text + "\n")
    # one HCA behind the same PCIe switch as the GPU (PIX), one on the other socket
    devs = {"ibp2": "pci0000:4f/0000:4f:01.0/0000:50:00.0/0000:51:00.0/0000:52:00.0",
            "ibp4": "pci0000:98/0000:98:01.0/0000:99:00.0/0000:9a:00.0/0000:9b:00.0"}
    gpu = "pci0000:4f/0000:4f:01.0/0000:50:00.0/0000:51:01.0/0000:53:00.0"
    for name, path in list(devs.items()) + [("gpu", gpu)]:
        d = f"{root}/devices/{path}"
        os.makedirs(d, exist_ok=True)
        bdf = path.split("/")[-1]
        w(f"{d}/numa_node", "0" if "4f" in path else "1")
        os.makedirs(f"{root}/bus/pci/devices", exist_ok=True)
        os.symlink(d, f"{root}/bus/pci/devices/{bdf}")
        if name != "gpu":
            ib = f"{d}/infiniband/{name}"
            w(f"{ib}/node_type", "1: CA")
            w(f"{ib}/ports/1/state", "4: ACTIVE")
            w(f"{ib}/ports/1/rate", "400 Gb/sec (4X NDR)")
            w(f"{ib}/ports/1/link_layer", "InfiniBand")
            w(f"{ib}/ports/1/lid", "0x12")
            os.makedirs(f"{root}/class/infiniband", exist_ok=True)
            os.symlink(ib, f"{root}/class/infiniband/{name}")
    w(f"{root}/module/nvidia_peermem/version", "580.159.03")


def test_probe_on_a_fake_gpu_box(tmp_path, monkeypatch):
    root = str(tmp_path / "sys")
    _fake_sysfs(root)
    monkeypatch.setenv("ROCNRDMA_SYSFS_ROOT", root)
    import importlib
    from rocnrdma_b200 import probe as P
    importlib.reload(P)
    hcas = P.list_hcas()
    assert [h.name for h in hcas] == ["ibp2", "ibp4"]
    assert hcas[0].active and hcas[0].rate_gbps == 400.0 and hcas[0].pci == "0000:52:00.0" and hcas[0].numa == 0
    g = P.Gpu(index=0, pci="0000:53:00.0", numa=0, pci_path=P._pci_path("0000:53:00.0"))
    assert P.pci_distance(g.pci_path, hcas[0].pci_path) == "PIX"
    assert P.pci_distance(g.pci_path, hcas[1].pci_path) == "SYS"
    assert P.affinity([g], hcas) == {0: "ibp2"}
    caps = P.capabilities()
    assert caps["has_hca_sysfs"] and caps["has_peermem"] and caps["peermem_version"] == "580.159.03"
    # exactly the GPU box of this project: HCAs in sysfs, no device nodes, no rdma-core -> software HCA
    caps["has_uverbs_dev"] = caps["has_libibverbs"] = caps["has_verbs"] = False
    plan = P.choose_wire(caps, n_gpus=1)
    assert plan["wire"] == "softhca" and "/dev/infiniband" in plan["why"]
    assert P.choose_wire(dict(caps, has_verbs=True), 1)["wire"] == "verbs"
    cfg = Cfg.Config().resolve({"plan": plan})
    assert cfg.wire == "softhca" and cfg.registration == "dmabuf"
    monkeypatch.delenv("ROCNRDMA_SYSFS_ROOT")
    importlib.reload(P)


def test_verbs_backend_reports_why_it_is_off():
    """Without the mock (a fresh process with ROCNRDMA_VERBS_LIBDIR unset) the backend looks for the system's
    rdma-core, finds none in this image, says so -- and `wire="auto"` falls back to the software HCA."""
    import os
    import subprocess
    import sys
    env = {k: v for k, v in os.environ.items() if k not in ("ROCNRDMA_VERBS_LIBDIR", "ROCNRDMA_WIRE")}
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from rocnrdma_b200 import _native as N, api\n"
            "lib = N.load()\n"
            "print(lib.rn_verbs_compiled(), lib.rn_verbs_available(), lib.rn_verbs_is_mock(), api.resolve_wire('auto'))\n"
            "print(lib.rn_verbs_why().decode())\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    first, why = out.stdout.strip().split("\n", 1)
    assert first == "1 0 0 softhca"                            # compiled in, nothing usable, not the mock
    assert "infiniband" in why.lower() or "ibverbs" in why.lower()
