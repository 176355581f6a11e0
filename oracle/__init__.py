"""CPU oracle for the sup3r GAN hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``sup3r_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker / the timed CPU stand-in.

PARITY STATUS: **parity unpinned** for the convolution / expansion arithmetic.
The reference (NREL/sup3r @ dd96e798) delegates all layer arithmetic to the
un-vendored third-party packages ``NREL-phygnn`` (pinned 0.0.33 in
``pixi.lock:320``) and ``tensorflow`` 2.15.1 / ``keras`` 2.15.0
(``pyproject.toml:54-57``), neither of which is installed or installable in the
build image, and the reference's own tests hold no golden vectors for any
generator output.  What IS pinned (see
``tests/test_oracle_pins.py``): the shape tables of
``tests/training/test_load_configs.py``, the loss identities of
``tests/utilities/test_loss_metrics.py``, the chunk index math fixtures of
``tests/forward_pass/test_forward_pass.py``, and an independent float64
cross-check of every layer against ``torch.nn.functional`` on CPU.
"""
