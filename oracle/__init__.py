"""CPU oracle for the dlgrad hot path — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  It holds
  kernels.c   a C restatement of the reference's generated kernels (each citing its source lines),
  build.py    compiles kernels.c with the reference's flags, and — when /root/reference is mounted —
              the reference's OWN generated C into oracle/_ref/ (git-ignored) for cross-checking,
  backend.py  registers (op, Device.CPU) functions in dlgrad_b200's dispatcher so that the same
              Tensor/ops layer can be run on the CPU restatement and compared with Device.CUDA.

Parity pinning: tests/golden/*.npz are outputs of the real reference (python package at
/root/reference, imported in the build container by tests/golden/make_golden.py); the oracle must
reproduce them (tests/test_oracle_golden.py) before any CUDA result is compared with it.
"""
