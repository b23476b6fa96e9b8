"""CPU oracle for the DR-NMF hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference algorithm (stwisdom/dr-nmf).
It exists to *check* the CUDA product path; it is never the thing shipped or measured
(except as the explicitly labelled ``cpu_baseline`` / ``--impl reference`` arm of bench.py).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.

Pinning status (see oracle/pin_reference.py and DESIGN.md):
  * The reference's own Python for ``SimpleDeepRNN.build/step``, ``build_alt``, ``DenseNonNegW.call``,
    ``DivideAbyAplusB``, ``ista_ed``, ``istft_noDiv``/``istft_mc``, ``stft_mc``, ``masked_seqs_to_frames`` and
    ``sparse_nmf_matlab``'s chunk driver was executed in the build container under a numpy stand-in for the
    Keras/Theano/librosa symbols it imports, and this oracle matches it (fixtures in tests/golden/).
  * Third-party semantics that are NOT in /root/reference (Keras 2.0.4 masked ``K.rnn`` scan, librosa 0.5.1
    ``stft`` conj convention, MATLAB ``mtimes`` order, BSS-Eval SDR) are restated from their published
    behaviour: for those pieces parity is UNPINNED.
"""
from .drnmf_oracle import *  # noqa: F401,F403
