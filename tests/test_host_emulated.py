"""Host logic without a GPU: the GPU parity suites (golden fixtures of the reference, oracle comparisons, error
behaviour, planners) run in a subprocess against tests/abi_emulator.py, a numpy stand-in for the C-ABI library.
This checks the Python host side and the host-side planners of the real library only; the CUDA kernels are checked
by the same suites with `-m gpu` on a B200."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gpu_suites_pass_on_the_abi_emulator():
    env = dict(os.environ, PYTHONPATH=HERE + os.pathsep + os.environ.get("PYTHONPATH", ""), T10_TEST_DEVICE="cpu")
    for k in ("TOR10_GGEMM_CFG", "TOR10_GGEMM_STREAMK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-p", "emu_plugin", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        os.path.join(HERE, "test_gpu_parity.py"), os.path.join(HERE, "test_gpu_configs.py")],
                       env=env, cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
