"""pytest plugin (`-p emu_plugin`, tests/ on PYTHONPATH): run the GPU parity suites on host memory against
tests/abi_emulator.py -- host-logic check for machines without a GPU.  See tests/test_host_emulated.py."""
import os
os.environ["T10_TEST_DEVICE"] = "cpu"
import abi_emulator  # noqa: F401,E402  (installs itself)
