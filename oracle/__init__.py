"""
oracle/ -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

A CPU (Python 3 + numpy/scipy) restatement of the per-iteration likelihood hot
path of irap-omp/deconv3d, used solely as the parity checker for the CUDA
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import anything from here; the
product package ``deconv3d_b200`` never does (``tests/test_no_oracle_in_product.py``
greps for it) and fails loudly when its CUDA library is missing.

Parity pinning (see DESIGN.md "Oracle"): the restatement in
``reference_port.py`` is checked bit-for-bit against chains produced by the
reference's own ``lib/run.py`` (executed in the build container with the
minimal py2->py3 source shims recorded in ``tests/golden/make_golden.py``) and
against the ``.mat`` known-answer fixture; ``rtnorm_port.py`` is checked
against the reference's ``lib/rtnorm.py`` fed with the same random draws.
"""
