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
``tests/forward_pass/test_forward_pass.py``, an independent float64
cross-check of every layer against ``torch.nn.functional`` on CPU, and the
known answers PUBLISHED in the TensorFlow / Keras API documentation of the ops
the networks are built from (``tf.nn.conv2d`` example -- cross-correlation,
(kh, kw, cin, cout) kernels, VALID --, the three ``tf.nn.depth_to_space``
examples, the ``tf.pad`` REFLECT / SYMMETRIC / CONSTANT examples,
``LeakyReLU``, ``sigmoid_cross_entropy_with_logits``, keras MSE / MAE): these
anchor the restated third-party algorithm on third-party vectors, they are not
outputs of the reference itself, so the status stays "unpinned" until a
``golden.npz`` exported from a real sup3r environment is dropped into
``tests/golden/phygnn/`` (``tools/export_phygnn_weights.py``).
"""
