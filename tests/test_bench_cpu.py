"""CPU: bench.py keeps stdout for its ONE JSON line -- library chatter written straight to file descriptor 1 (NCCL's
version banner is a printf to stdout) must come out on stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stray_stdout_goes_to_stderr():
    code = ("import os, sys; sys.path.insert(0, %r); import bench; "
            "bench._dispatch = lambda a: (os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'), print('{\"ok\": 1}')); "
            "sys.argv = ['bench.py']; bench.main()" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().splitlines() == ['{"ok": 1}'] and json.loads(r.stdout) == {"ok": 1}
    assert "NCCL version" in r.stderr
