"""CPU oracle for the OSRL per-step training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline -- never as the thing shipped.  The product path
(``osrl_b200``) fails loudly when its CUDA library is missing; it never routes
through this package.

What it is: a plain-PyTorch fp32 (CPU) functional restatement of the reference's
``train_one_step`` for BC, BCQ-Lag, CPQ, BEAR-Lag and CDT plus the two minibatch
samplers, written against the reference lines cited in each function
(paths relative to /root/reference).  The arithmetic of the reference lives in
PyTorch itself (``nn.Linear``, autograd, ``torch.optim.Adam``), which is a
third-party dependency pinned ``torch~=1.13`` in the reference's setup.py:25 and
present here as torch 2.11; the restatement spells out Adam, Polyak, the PID
controller and every loss explicitly so the CUDA kernels have a line-by-line
specification.

Parity pin: the reference ships no tests or golden vectors, so the
pin is the reference itself executed in the build container:
``oracle/make_golden.py`` imports the unmodified reference from /root/reference
(with stub ``gymnasium``/``fsrl`` modules), runs both on identical seeds, batches
and replayed noise, asserts bit-for-bit (or <=1e-6 relative where summation order
differs) agreement, and writes the fixtures under ``tests/golden/``.
"""
