"""Parity oracle for the PULPo hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``pulpo_b200``) never
does; it fails loudly when its CUDA library is missing.

* ``oracle.cport``      ctypes binding of ``pulpo_oracle.c`` (plain C, fp32, op-by-op rounding)
* ``oracle.torch_ref``  restatement over torch-CPU ATen ops (the reference's own arithmetic
                        lives in ATen: grid_sampler_3d, upsample_trilinear3d, conv3d)
* ``oracle.ref_import`` loads the real reference from /root/reference (build container only)
* ``oracle.gen_golden`` writes tests/golden/*.npz from the real reference
"""
