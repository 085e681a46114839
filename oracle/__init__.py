"""CPU oracle for the deephar forward hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's forward path
(deephar/models/reception.py, models/spnet.py, models/common.py, models/blocks.py,
layers.py, activations.py) together with the Keras-2.1.4 / TF-1.6 op semantics
those files rely on.

PARITY UNPINNED: the reference's arithmetic lives in keras==2.1.4 and
tensorflow-gpu==1.6.0 (requirements.txt:2-3), neither of which is importable in
this image, the reference ships no tests or golden tensors, and its released
weights are download-only.  The only piece of reference source that can be
executed here is deephar/utils/math.py::linspace_2d (pure numpy); the fixture
tests/golden/linspace_2d.npz was produced by executing that source text
(tests/golden/make_golden.py) and pins the soft-argmax grid.  Everything else is
a restatement checked by closed-form known-answer tests (tests/test_oracle_kat.py)
and by an independent torch-CPU implementation of the same ops (ops_torch.py).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this package.  The product (deephar_b200/) never does.
"""
