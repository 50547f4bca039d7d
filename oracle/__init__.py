"""CPU oracle for the video-prediction training step  --  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the algorithm in the reference's
``src/models/train_model.py:42-764`` (model, helpers) and ``:860-861,950-960``
(Adam step) together with the Chainer 2.0.1 operator semantics that file relies
on (``chainer==2.0.1`` is pinned in the reference's ``requirements.txt:10`` and is
NOT vendored in ``/root/reference`` nor installable here).

PARITY UNPINNED: the reference holds no golden vectors, known-answer tests or
fixtures for this path, and neither Chainer nor Python 2 can run in this image,
so the oracle cannot be checked against the reference's own outputs.  It is
pinned instead by (i) an independent torch-CPU autograd restatement
(``tests/torch_restatement.py``), (ii) float64 finite differences and
(iii) frozen vectors under ``tests/golden/`` produced by ``tests/golden/make_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline.  The product path (``physical-interaction-video-prediction_b200``)
never imports it and has no CPU fallback.
"""
