"""dr-nmf_b200: B200-native (sm_100a) implementation of the DR-NMF hot path of stwisdom/dr-nmf.

Layout
  csrc/            CUDA kernels + the C-ABI (built in-tree into libdrnmf.so by csrc/build.sh)
  _lib.py          ctypes binding of include/drnmf.h (fails loudly if the library or a B200 is missing)
  engine.py        thin torch-tensor wrapper of the C-ABI (device memory, streams; no compute in Python)
  custom_layers.py / enhance.py / snmf.py / util.py / audio_dataset.py
                   host-side mirrors of the reference's interfaces for this path (same names and argument meaning)
  synth.py         synthetic CHiME2-shaped workload generator

Import as `drnmf_b200` (see ../drnmf_b200/__init__.py).
"""
__version__ = "0.1.0"
