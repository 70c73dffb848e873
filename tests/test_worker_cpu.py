"""Persistent stage worker (acezero_b200/worker.py): a stage forwarded to the worker runs in the worker's
long-lived process with the client's argv and working directory, its output and exit status come back through the client, a failed
stage does not take the worker down, and without a worker the client falls through to in-process execution."""
import os
import subprocess
import sys
import textwrap
import time
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

STAGE = textwrap.dedent('''
    import os, sys
    def main(argv=None):
        print("stage pid", os.getpid(), "cwd", os.getcwd(), "argv", argv, flush=True)
        print("to stderr", file=sys.stderr, flush=True)
        if "--write" in argv:
            open("out.txt", "w").write("written by the stage")
        if "--fail" in argv:
            raise RuntimeError("stage failed")
        if "--exit3" in argv:
            sys.exit(3)
''')

CLIENT = textwrap.dedent('''
    import sys
    sys.path.insert(0, {root!r})
    from acezero_b200 import worker
    worker.try_forward("acez_dummy_stage", sys.argv[1:])
    print("ran locally", flush=True)
''')


@pytest.fixture
def served(tmp_path):
    (tmp_path / "acez_dummy_stage.py").write_text(STAGE)
    (tmp_path / "client.py").write_text(CLIENT.format(root=str(ROOT)))
    sock = str(tmp_path / "w.sock")
    env = dict(os.environ, PYTHONPATH=f"{tmp_path}{os.pathsep}{ROOT}", ACEZ_WORKER_EXTRA_ENTRIES="acez_dummy_stage")
    env.pop("ACEZ_WORKER", None)
    srv = subprocess.Popen([sys.executable, "-m", "acezero_b200.worker", "serve", "--socket", sock, "--no-warmup"], cwd=str(ROOT), env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    for _ in range(100):
        if os.path.exists(sock):
            break
        time.sleep(0.1)
    assert os.path.exists(sock), "worker did not start"
    yield tmp_path, sock, env, srv
    subprocess.run([sys.executable, "-m", "acezero_b200.worker", "shutdown", "--socket", sock], cwd=str(ROOT), env=env, timeout=30)
    try:
        srv.wait(timeout=10)
    except subprocess.TimeoutExpired:
        srv.kill()


def _client(tmp, env, sock, *args, cwd=None):
    e = dict(env, ACEZ_WORKER=sock)
    return subprocess.run([sys.executable, str(tmp / "client.py"), *args], cwd=str(cwd or tmp), env=e, capture_output=True, text=True, timeout=60)


def test_stages_run_in_the_persistent_process(served):
    tmp, sock, env, srv = served
    work = tmp / "work"
    work.mkdir()
    a = _client(tmp, env, sock, "--write", "x", cwd=work)
    b = _client(tmp, env, sock, "second")
    assert a.returncode == 0 and b.returncode == 0, (a.stdout, a.stderr, b.stdout)
    assert "ran locally" not in a.stdout + b.stdout
    pid = lambda out: out.split("stage pid ")[1].split()[0]
    assert pid(a.stdout) == pid(b.stdout) == str(srv.pid)                 # both stages ran in the worker, not in the clients
    assert f"cwd {work}" in a.stdout and "'--write', 'x'" in a.stdout     # the client's working directory and arguments
    assert "to stderr" in a.stdout                                        # stderr merged into the stream (ace_zero_util.py:33)
    assert (work / "out.txt").read_text() == "written by the stage"


def test_exit_status_and_survival(served):
    tmp, sock, env, srv = served
    f = _client(tmp, env, sock, "--fail")
    assert f.returncode == 1 and "RuntimeError: stage failed" in f.stdout   # traceback reaches the caller, status is non-zero
    e = _client(tmp, env, sock, "--exit3")
    assert e.returncode == 3
    ok = _client(tmp, env, sock, "after")
    assert ok.returncode == 0 and str(srv.pid) in ok.stdout               # the worker survived both


def test_without_a_worker_the_stage_runs_locally(tmp_path):
    (tmp_path / "acez_dummy_stage.py").write_text(STAGE)
    (tmp_path / "client.py").write_text(CLIENT.format(root=str(ROOT)))
    env = dict(os.environ, PYTHONPATH=f"{tmp_path}{os.pathsep}{ROOT}", ACEZ_WORKER=str(tmp_path / "nobody.sock"))
    r = subprocess.run([sys.executable, str(tmp_path / "client.py"), "x"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ran locally" in r.stdout
    env.pop("ACEZ_WORKER")
    r = subprocess.run([sys.executable, str(tmp_path / "client.py"), "x"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ran locally" in r.stdout


def test_stage_clis_do_not_forward_multi_gpu_or_ranked_jobs(monkeypatch):
    from acezero_b200 import worker
    monkeypatch.setenv("ACEZ_WORKER", "/nonexistent.sock")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "2")
    assert worker.try_forward("train_ace", []) is None                     # inside torchrun: never forwarded
    monkeypatch.delenv("RANK")
    monkeypatch.delenv("WORLD_SIZE")
    monkeypatch.setenv("ACEZ_IN_WORKER", "1")
    assert worker.try_forward("train_ace", []) is None                     # inside the worker itself
