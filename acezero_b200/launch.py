"""Process placement for the two stage executables (`train_ace.py`, `register_mapping.py`).

`ace_zero.py` starts every stage as a plain subprocess (reference ace_zero_util.py:8-9,32: `subprocess.Popen` of
`./train_ace.py ...`) and must stay unchanged, so the multi-GPU plumbing lives INSIDE the executables:

  * full stages (mapping of an iteration, registration of all images): when `ACEZ_GPUS=G` (or `--gpus G`) asks for
    G > 1 GPUs, the executable replaces itself with
        python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port <free> <script> <same argv>
    and every rank (one process per GPU, NCCL) re-enters `main()` with RANK / LOCAL_RANK / WORLD_SIZE set. Exit status and
    the stdout/stderr stream are those of the torchrun process, which is what ace_zero_util.run_cmd consumes (:32-49).
  * small jobs — the seed trials of ace_zero.py:184-196 (`--use_pose_seed >= 0`: one mapping image) and their fast
    registration check (`--max_estimates > 0`) — stay on ONE GPU each, and several of them started in parallel by
    `--seed_parallel_workers` (joblib, ace_zero.py:195) spread over the box's GPUs through an advisory lock per device
    ("one seed per GPU"; no collective, results travel through files as in the reference).
"""
import fcntl
import os
import socket
import sys


def requested_gpus(flag_value=0):
    """Number of ranks a full stage should use: --gpus N > ACEZ_GPUS (integer or 'all') > 1."""
    if flag_value and int(flag_value) > 0:
        return int(flag_value)
    v = os.environ.get("ACEZ_GPUS", "").strip().lower()
    if not v:
        return 1
    if v == "all":
        import torch
        return max(1, torch.cuda.device_count())
    return max(1, int(v))


def in_worker():
    """True inside a rank started by torchrun (RANK / WORLD_SIZE in the environment)."""
    return "RANK" in os.environ and "WORLD_SIZE" in os.environ


def world():
    """(rank, world_size, local_rank) of this process; (0, 1, 0) outside torchrun."""
    if not in_worker():
        return 0, 1, 0
    return int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))


def free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun_command(script, argv, n_gpus, port=None):
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={int(n_gpus)}",
            "--master-addr", "127.0.0.1", "--master-port", str(port if port is not None else free_port()),
            str(script)] + [str(a) for a in argv]


def maybe_self_launch(script, argv, n_gpus, small_job=False):
    """Replace this process by a torchrun of itself when more than one GPU is requested. Returns only when the caller
    should do the work itself (single GPU, or already a rank of the group)."""
    if in_worker() or n_gpus <= 1 or small_job:
        return
    cmd = torchrun_command(script, argv, n_gpus)
    sys.stdout.flush()
    sys.stderr.flush()
    os.execv(sys.executable, cmd)  # same pid: exit status and output stream reach ace_zero_util.run_cmd unchanged


def init_distributed():
    """Join the process group of this torchrun (NCCL on GPUs) and select the rank's device. Returns (rank, world)."""
    import torch
    rank, ws, local = world()
    if ws == 1:
        return 0, 1
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, ws


def shutdown_distributed():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


_lease_fd = None


def lease_gpu(n_devices, lock_dir=None, blocking_fallback=True):
    """Pick a GPU for a single-GPU process so that concurrent stage processes of one box (parallel seed trials) do not pile
    up on device 0: the first device whose advisory lock (flock on <lock_dir>/gpu<i>.lock, held for the life of the
    process) is free. When every device is taken: wait for device (pid mod n) if `blocking_fallback`, else share it.
    Returns the device index. An explicit CUDA_VISIBLE_DEVICES with a single device needs no lease."""
    global _lease_fd
    if n_devices <= 1:
        return 0
    d = lock_dir or os.path.join(os.environ.get("TMPDIR", "/tmp"), f"acez_gpu_lease_{os.getuid()}")
    os.makedirs(d, exist_ok=True)
    for i in range(n_devices):
        fd = os.open(os.path.join(d, f"gpu{i}.lock"), os.O_CREAT | os.O_RDWR, 0o600)
        try:
            fcntl.flock(fd, fcntl.LOCK_EX | fcntl.LOCK_NB)
            _lease_fd = fd
            return i
        except OSError:
            os.close(fd)
    i = os.getpid() % n_devices
    if blocking_fallback:
        fd = os.open(os.path.join(d, f"gpu{i}.lock"), os.O_CREAT | os.O_RDWR, 0o600)
        fcntl.flock(fd, fcntl.LOCK_EX)
        _lease_fd = fd
    return i


def release_gpu():
    global _lease_fd
    if _lease_fd is not None:
        try:
            fcntl.flock(_lease_fd, fcntl.LOCK_UN)
            os.close(_lease_fd)
        finally:
            _lease_fd = None


def select_device(small_job=False):
    """Device set-up of a stage executable. Inside torchrun: the rank's GPU + process group. Otherwise one GPU, leased
    when the box has several (so parallel seed workers land on different GPUs). Returns (rank, world_size)."""
    import torch
    if in_worker():
        return init_distributed()
    n = torch.cuda.device_count()
    if n > 1 and os.environ.get("ACEZ_GPU_LEASE", "1") != "0":
        torch.cuda.set_device(lease_gpu(n))
    return 0, 1
