"""Shared start-up of the multi-GPU scripts (tests/mgpu_*.py, run under torchrun with one rank per GPU).
MGPU_DRY=1 checks a script's logic on the CPU: one gloo rank, the oracle behind the same API (what the CPU suite runs)."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402


def setup():
    """-> (backend, local device, rank, world, torch device for host<->rank gathers)"""
    import torch
    import torch.distributed as dist
    if os.environ.get("MGPU_DRY") == "1":
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        return vh.load_backend(os.path.join(ROOT, "oracle", "_build", "libvahana_oracle.so")), 0, 0, 1, torch.device("cpu")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = vh.default_backend()
    be.init(local)
    be.set_stream(torch.cuda.current_stream().cuda_stream)
    be.init_distributed()
    return be, local, dist.get_rank(), dist.get_world_size(), torch.device("cuda", local)


def oracle_backend():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return vh.load_backend(os.path.join(ROOT, "oracle", "_build", "libvahana_oracle.so"))


def run_ranks(script, port, nranks=None, env=None, timeout=600):
    """Run tests/<script> under torchrun with one rank per visible GPU (at most 4); on failure the assertion message carries the
    ranks' own tracebacks only (torchrun's elastic report is dropped)."""
    import subprocess
    import torch
    ng = torch.cuda.device_count() if nranks is None else nranks
    n = min(ng, 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", script)], capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, **(env or {})))
    if r.returncode != 0 or r.stdout.count(": ok") != n:
        keep = [ln for ln in r.stderr.splitlines() if ln.startswith("[rank") or "Error" in ln and "elastic" not in ln]
        raise AssertionError(f"{script} on {n} ranks: rc {r.returncode}\n" + r.stdout[-1500:] + "\n" + "\n".join(keep[-40:]))
    return r.stdout
