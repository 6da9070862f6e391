"""CPU oracle for the ConMamba hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic the reference (mattmireles/Mamba-ASR)
runs on the hot path named in BASELINE.json:

  * ``scan_ref``    - ``selective_scan_ref``            (modules/mamba/selective_scan_interface.py:91-157)
  * ``conv_ref``    - the torch causal-conv fallback     (modules/mamba/bimamba.py:83-91, 278-279)
  * ``bimamba_ref`` - the BiMamba-v2 composition         (modules/mamba/bimamba.py:176-253)
  * ``fbank_ref``   - SpeechBrain 1.0.0 ``Fbank``        (called at train_CTC.py:285; arithmetic lives in
                                                          speechbrain==1.0.0, requirement.txt:11)
  * ``lengths_ref`` - length / padding-mask integers     (modules/TransformerASR.py:408-410)

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker.  Nothing under ``mamba_asr_b200/`` imports
``oracle``; the product path has no CPU fallback and raises if the CUDA library is missing.

Pinning status
--------------
* scan     : PINNED.  ``tests/golden/make_golden.py`` imports the reference's own
             ``selective_scan_ref`` (module stubs for the three absent CUDA extensions) in the build
             container and freezes its outputs and autograd gradients in ``tests/golden/scan_*.npz``;
             ``tests/test_oracle.py`` checks this restatement against them bit-for-bit (fp32).
* conv     : pinned to the reference's own torch fallback expression (``bimamba.py:278-279``) evaluated by
             ``make_golden.py`` with an ``nn.Conv1d`` built exactly as ``bimamba.py:83-91``.
* bimamba  : pinned by composing the two pinned pieces literally as ``bimamba.py:223-253`` does.
* fbank    : PARITY UNPINNED.  speechbrain is not vendored in the reference and not installable here;
             the restatement follows the recalled 1.0.0 source and is only
             cross-checked against an independent numpy DFT implementation.
"""
