"""CPU oracle for the EYOC registration-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``eyoc_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker (or as the timed CPU
baseline), never as part of the shipped path.

Parity status:
  * SC2-PCR / kNN / Kabsch / SE3 / IRLS restatements are PINNED: they are
    bit-compared against the reference's own functions imported from
    /root/reference by ``oracle/pin_against_reference.py``, which also writes
    the committed fixtures under ``tests/golden/``.
  * The ResUNetBN2C restatement is "parity unpinned" against MinkowskiEngine
    itself: ME is an un-vendored, un-pinned third-party dependency that is
    absent from /root/reference and from this image.  It is pinned instead
    against dense ``torch.nn.functional.conv3d`` / ``conv_transpose3d`` on a
    densified grid (self-consistency of the published ME algorithm).
"""
