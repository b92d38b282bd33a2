"""CPU oracle for the MSF-WSI SSL head + loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product path (``msfwsi_b200``) never imports this
package and has no CPU fallback.

Parity status (see DESIGN.md "Oracle"):
  * reference-pinned: crop coordinates, jigsaw indices, inverse gather, fuser
    concat, projector / predictor heads, SimSiam negative-cosine loss block.
    Pinned against outputs of the unmodified reference module imported from
    /root/reference (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
  * parity unpinned (the reference has no such code and no tests): InfoNCE
    mode, bilinear crop-resample (only its integer case is pinned, against
    hooknet.py:29-32), EMA update.  Their oracle is the fp64 restatement here.
"""
