"""CPU oracle for the tactileSR hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch-CPU tensor arithmetic, fp32 or fp64)
of the reference algorithm in ``model/tactileSR_model.py``, ``model/tPSFNet.py``
and the ``train_cal_loss`` / Adam step around them.  It is the *checker* for the
CUDA path in ``tactilesr_b200``; it is never the thing that is shipped or
measured as the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  Nothing under ``tactilesr_b200/``
imports it, and the product path raises if the CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the *unmodified reference modules*
imported from ``/root/reference`` in the build container; the generating
script is ``oracle/make_golden.py`` and the vectors are ``tests/golden/*.npz``.
"""
