"""CPU oracle for the TensorLy-Torch factorized-layer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``torch_b200/`` (the product) may import,
call, link or execute anything in this package.  The only permitted users are
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- and there only as the checker / the CPU baseline, never as the
thing measured or shipped.

What is in here
---------------
``tensorly_semantics``  restatement of the third-party ``tensorly`` functions the reference
                        calls on this path (tensorly is NOT vendored under /root/reference, is
                        un-pinned in its requirements.txt:5 and is not installable offline).
``reference_path``      restatement, in plain torch on plain tensors, of the reference's
                        ``tltorch/functional/*.py``, the ``to_tensor`` reconstructions and the
                        tensor-dropout hooks.  Every function cites the file:line it follows.
``_shim/tensorly``      import shim (adapter over ``tensorly_semantics``) that lets the REAL
                        reference package be imported in the build container so that
                        ``tests/golden/make_golden.py`` can run the reference's own code and pin
                        the restatement against it.

Parity status: the reference-owned code (einsum-string construction, conv chains, BlockTT
reconstruction, dropout hooks) is PINNED against outputs of the real reference executed here
(tests/golden/*.pt).  The tensorly layer underneath is pinned only by the textbook definition of
each function and by the reference's own self-consistency identities (factorized path == dense op
on the reconstructed weight), because no tensorly source or wheel is available offline.
"""
