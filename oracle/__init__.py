"""CPU oracle for the AVR acoustic volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``avr_b200/`` may import this package; the
only legal importers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Contents
--------
render_ref.py     fp32 torch restatement of the reference renderer
                  (``/root/reference/renderer_cpu.py:23-171``), written from the maths in
                  SURVEY.md Appendix A.  PINNED: checked bit-for-bit / to 1e-6 against the
                  unmodified reference module in ``tests/test_oracle_vs_reference.py`` (runs
                  wherever ``/root/reference`` exists) and against the committed golden
                  vectors in ``tests/golden/`` (runs everywhere).
field_ref.py      fp32 torch restatement of the tiny-cuda-nn pieces the reference's
                  ``model.py`` instantiates (HashGrid encoding, bias-free MLPs).
                  PARITY UNPINNED: tiny-cuda-nn is an un-vendored, un-pinned dependency
                  (``/root/reference/requirements.txt:11``) that is neither installed nor
                  installable here, and the reference holds no tests or golden vectors for
                  it; the restatement follows the published algorithm.
criterion_ref.py  fp32 torch restatement of the training loss (``utils/criterion.py:7-126``) and of
                  the ``auraloss`` multi-resolution STFT loss it calls (absent, unpinned: that term is
                  PARITY UNPINNED; everything else is PINNED against the unmodified reference module).
reference_shim.py imports ``/root/reference/renderer_cpu.py``, ``model.py`` (with the oracle's HashGrid /
                  MLP standing in for ``tinycudann``) and ``utils/criterion.py`` (stand-in ``auraloss``)
                  unmodified (this container only; the GPU box has no ``/root/reference``).
make_golden.py    regenerates ``tests/golden/*.npz`` from the unmodified reference renderer
                  driven by ``field_ref`` networks.
"""
