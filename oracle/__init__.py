"""CPU oracle for the deephar forward hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's forward path
(deephar/models/reception.py, models/spnet.py, models/common.py, models/blocks.py,
layers.py, activations.py) together with the Keras-2.1.4 / TF-1.6 op semantics
those files rely on.

PARITY -- what is pinned by the reference itself and what is not:

* GRAPH LEVEL: PINNED.  tests/golden/make_reference_golden.py imports the reference's own builder code
  (deephar/models/reception.py, spnet.py, action.py, blocks.py, common.py, layers.py, activations.py --
  unmodified, from /root/reference) and EXECUTES it on tests/golden/keras_shim, an eager float64 stand-in
  for the Keras 2.1.4 functional API.  The fixtures tests/golden/ref_*.npz (ReceptionNet 2-D ctx / heat-map
  export / 3-D / full-size BASELINE configs[1] model, SPNet Penn-like / NTU-like / pose-only / full-resolution
  configs[3] architecture, CVPR'18 merge model) hold the reference models'
  weight lists (Keras auto-names, shapes) and outputs; tests/test_reference_golden.py requires
  reception.py / spnet.py / action.py here to reproduce them to 1e-9 (fp64) and the product's weight_specs
  to be exactly the reference's learned weights.
  Also pinned: the soft-argmax grid (tests/golden/linspace_2d.npz, produced by executing
  deephar/utils/math.py::linspace_2d, the one numeric function of the reference that runs without Keras).
* PRIMITIVE OP SEMANTICS: UNPINNED.  What a Keras/TF op computes (TF 'SAME' padding, BatchNormalization
  inference formula and epsilon, max-pool padding value, SeparableConv2D composition, ...) lives in
  keras==2.1.4 / tensorflow-gpu==1.6.0 (requirements.txt:2-3), neither importable in this image; the
  reference ships no tests or golden tensors and its weights are download-only.  ops_np.py (and the shim,
  written independently with torch kernels) restate SURVEY.md Appendix A; they are checked by closed-form
  known-answer tests (tests/test_oracle_kat.py) and against each other, not against TensorFlow.
  Third-party referees that ARE in the image (tests/test_oracle_thirdparty.py): the TF-'SAME' padding rule against
  HuggingFace transformers' TensorFlow-compatible padding (apply_tf_padding of its TF-ported MobileNets), and the
  arithmetic of Conv2D / the SeparableConv2D stages / MaxPooling2D (-inf padding) / AveragePooling2D /
  BatchNormalization(eps 1e-3) / UpSampling2D / soft-max against torch's own kernels behind that padding.  That Keras
  2.1.4 maps its arguments onto exactly these rules remains a documented fact, not an executed one.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this package.  The product (deephar_b200/) never does.
"""
