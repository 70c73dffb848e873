"""Persistent stage worker (reference ace_zero_util.py:8-52).

`ace_zero.py` starts every mapping / registration stage as a fresh subprocess (`subprocess.Popen(["./train_ace.py", ...])`,
ace_zero_util.py:32), so every stage pays the interpreter start, `import torch`, the CUDA context, the load of libacez.so and
the first-launch set-up of its kernels again - seconds per stage, dozens of stages per reconstruction, while the stage itself
takes seconds on this GPU path. The outer loop must stay unchanged, so the persistence lives behind the executables:

    python -m acezero_b200.worker serve [--socket PATH]     # once, before ace_zero.py (or ACEZ_WORKER=auto: started on first use)
    ACEZ_WORKER=1 ./ace_zero.py ...                         # ./train_ace.py / ./register_mapping.py forward their argv

With `ACEZ_WORKER` set, a stage executable parses its arguments as always and then hands (entry, argv, cwd) to the worker over a
Unix socket; the worker runs the SAME `main(argv)` in its long-lived process with stdout / stderr redirected into the connection
(the reference merges them anyway, ace_zero_util.py:33) and returns the exit status, which the thin client re-raises - so
`run_cmd` sees the same stream and the same return code (non-zero aborts ACE0, :48-49). Jobs are served one at a time (one GPU,
one process; parallel seed trials queue up). Multi-GPU stages (`ACEZ_GPUS` / `--gpus` > 1) are never forwarded: they re-launch
themselves under torchrun. Without `ACEZ_WORKER`, or when no worker answers, the executable runs in its own process as before.

Only the standard library is imported here: the client path must stay cheap.
"""
import json
import os
import socket
import subprocess
import sys
import tempfile
import time
import traceback

SENTINEL = b"\x00ACEZ_EXIT "
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENTRIES = ("train_ace", "register_mapping", "export_point_cloud")


def default_socket():
    return os.environ.get("ACEZ_WORKER_SOCKET") or os.path.join(tempfile.gettempdir(), f"acez_worker_{os.getuid()}.sock")


def _allowed_entries():
    extra = [e.strip() for e in os.environ.get("ACEZ_WORKER_EXTRA_ENTRIES", "").split(",") if e.strip()]
    return set(ENTRIES) | set(extra)


# ----------------------------------------------------------------------------------------------------------------
# server
# ----------------------------------------------------------------------------------------------------------------
def _run_entry(entry, argv):
    import importlib
    if entry not in _allowed_entries():
        raise ValueError(f"unknown stage entry '{entry}'")
    mod = importlib.import_module(entry)
    rc = mod.main(list(argv))
    return 0 if rc is None else int(rc)


def _handle(conn, state):
    f = conn.makefile("rb")
    line = f.readline()
    if not line:
        return True
    req = json.loads(line.decode())
    op = req.get("op", "run")
    if op == "ping":
        conn.sendall(json.dumps({"pid": os.getpid(), "jobs": state["jobs"]}).encode() + b"\n")
        return True
    if op == "shutdown":
        conn.sendall(b"bye\n")
        return False
    state["jobs"] += 1
    sys.stdout.flush()
    sys.stderr.flush()
    saved = (os.dup(1), os.dup(2))
    cwd0 = os.getcwd()
    code = 1
    try:
        os.dup2(conn.fileno(), 1)
        os.dup2(conn.fileno(), 2)
        try:
            if req.get("cwd"):
                os.chdir(req["cwd"])
            code = _run_entry(req["entry"], req.get("argv", []))
        except SystemExit as e:   # argparse errors, explicit exits
            code = e.code if isinstance(e.code, int) else (0 if e.code is None else 1)
            if not isinstance(e.code, int) and e.code is not None:
                print(e.code, file=sys.stderr)
        except BaseException:  # noqa: BLE001  (a failed stage must not take the worker down; the client gets status 1)
            traceback.print_exc()
            code = 1
    finally:
        try:
            sys.stdout.flush()
            sys.stderr.flush()
        finally:
            os.dup2(saved[0], 1)
            os.dup2(saved[1], 2)
            os.close(saved[0])
            os.close(saved[1])
            os.chdir(cwd0)
    conn.sendall(SENTINEL + str(int(code)).encode() + b"\n")
    return True


def serve(socket_path=None, warmup=True, idle_timeout=0.0):
    """Serve stage jobs until a shutdown request (or `idle_timeout` seconds without one, 0 = forever)."""
    path = socket_path or default_socket()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["ACEZ_IN_WORKER"] = "1"
    os.environ["ACEZ_GPUS"] = "1"        # a forwarded stage never re-launches itself under torchrun (that would replace the worker)
    os.environ.pop("ACEZ_WORKER", None)
    if os.path.exists(path):
        if _connect(path) is not None:
            raise RuntimeError(f"a worker already serves {path}")
        os.unlink(path)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    os.chmod(path, 0o600)
    srv.listen(64)
    if warmup:
        # what every stage would otherwise pay again: torch, the CUDA context, the library
        try:
            import torch
            if torch.cuda.is_available():
                torch.zeros(1, device="cuda")
            from . import _lib
            _lib.load()
        except Exception:  # noqa: BLE001  (a worker without a GPU still serves; the stage itself will fail loudly)
            traceback.print_exc()
    state = {"jobs": 0}
    try:
        while True:
            srv.settimeout(idle_timeout if idle_timeout and idle_timeout > 0 else None)
            try:
                conn, _ = srv.accept()
            except socket.timeout:
                break
            with conn:
                conn.settimeout(None)
                if not _handle(conn, state):
                    break
    finally:
        srv.close()
        try:
            os.unlink(path)
        except OSError:
            pass


# ----------------------------------------------------------------------------------------------------------------
# client
# ----------------------------------------------------------------------------------------------------------------
def _connect(path, timeout=2.0):
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(timeout)
    try:
        s.connect(path)
    except OSError:
        s.close()
        return None
    s.settimeout(None)
    return s


def _spawn(path):
    """Start a detached worker for `path` and wait until it listens (torch import + CUDA context: up to a few minutes cold)."""
    log = open(path + ".log", "ab")
    subprocess.Popen([sys.executable, "-m", "acezero_b200.worker", "serve", "--socket", path, "--idle-timeout",
                      os.environ.get("ACEZ_WORKER_IDLE", "1800")], cwd=ROOT, stdout=log, stderr=log, stdin=subprocess.DEVNULL,
                     start_new_session=True, env={k: v for k, v in os.environ.items() if k != "ACEZ_WORKER"})
    t0 = time.time()
    while time.time() - t0 < float(os.environ.get("ACEZ_WORKER_START_TIMEOUT", "300")):
        c = _connect(path)
        if c is not None:
            return c
        time.sleep(0.2)
    return None


def forward(entry, argv, path=None, out=None):
    """Run `entry` with `argv` in the worker behind `path`; streams its output to `out` (default: this process's stdout).
    Returns the exit status, or None when no worker answered."""
    path = path or default_socket()
    conn = _connect(path)
    if conn is None and os.environ.get("ACEZ_WORKER", "").strip().lower() == "auto":
        conn = _spawn(path)
    if conn is None:
        return None
    out = out if out is not None else getattr(sys.stdout, "buffer", sys.stdout)
    with conn:
        conn.sendall(json.dumps({"op": "run", "entry": entry, "argv": [str(a) for a in argv], "cwd": os.getcwd()}).encode() + b"\n")
        tail = b""
        keep = len(SENTINEL) + 16
        while True:
            chunk = conn.recv(65536)
            if not chunk:
                break
            tail += chunk
            if len(tail) > keep:   # everything but the last `keep` bytes cannot belong to the trailer
                out.write(tail[:-keep])
                out.flush()
                tail = tail[-keep:]
        k = tail.rfind(SENTINEL)
        if k < 0:                  # the worker died mid-job
            out.write(tail)
            out.flush()
            return 1
        out.write(tail[:k])
        out.flush()
        try:
            return int(tail[k + len(SENTINEL):].split(b"\n")[0])
        except ValueError:
            return 1


def try_forward(entry, argv):
    """Called by the stage executables after argument parsing, single-GPU jobs only. No-op unless ACEZ_WORKER is set
    (1 / auto: the default socket, anything else: a socket path); exits the process with the job's status when a worker ran it."""
    mode = os.environ.get("ACEZ_WORKER", "").strip()
    if not mode or mode == "0" or os.environ.get("ACEZ_IN_WORKER") or ("RANK" in os.environ and "WORLD_SIZE" in os.environ):
        return
    path = default_socket() if mode.lower() in ("1", "auto", "true", "yes") else mode
    code = forward(entry, argv, path)
    if code is None:
        return            # nobody there: run in this process, as without the switch
    sys.exit(code)


def _main():
    import argparse
    ap = argparse.ArgumentParser(description="persistent stage worker of acezero-b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("serve")
    s.add_argument("--socket", default=None)
    s.add_argument("--no-warmup", action="store_true")
    s.add_argument("--idle-timeout", type=float, default=0.0)
    for name in ("ping", "shutdown"):
        p = sub.add_parser(name)
        p.add_argument("--socket", default=None)
    a = ap.parse_args()
    if a.cmd == "serve":
        serve(a.socket, warmup=not a.no_warmup, idle_timeout=a.idle_timeout)
        return 0
    c = _connect(a.socket or default_socket())
    if c is None:
        print("no worker")
        return 1
    with c:
        c.sendall(json.dumps({"op": a.cmd}).encode() + b"\n")
        print(c.makefile("rb").readline().decode().strip())
    return 0


if __name__ == "__main__":
    sys.exit(_main())
