"""CPU oracle for the EAS-SNN hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy / plain PyTorch on the CPU, the algorithm of the
reference's hot path so that the CUDA kernels in ``eas_snn_b200`` can be checked
against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``eas_snn_b200/`` imports it, and the product path raises if the CUDA library is
missing instead of falling back to this code.

Parity pinning
--------------
* ``oracle.binning``  follows ``yolox/data/datasets/gen1.py:313-360``            -- PINNED
  by golden vectors generated from the reference itself (``tests/golden/make_golden.py``).
* ``oracle.sampler``  follows ``yolox/models/embedding.py:132-226`` and
  ``yolox/models/activation.py:17-30``                                           -- PINNED
  (forward and parameter gradients, golden vectors from the reference).
* ``oracle.plif``     restates ``spikingjelly==0.0.0.0.14``
  ``activation_based/neuron.py`` (``ParametricLIFNode``), ``surrogate.py`` (``ATan``,
  ``Sigmoid``), which is a third-party dependency NOT vendored under the reference
  tree (``pip-requirements.txt:135``) and not installed here               -- PARITY UNPINNED
  at the spikingjelly boundary (no reference test or fixture covers it); the
  restatement is pinned only structurally (call sites ``yolox/utils/utils_snn.py:44-53``).
* ``oracle.backbone`` follows ``yolox/models/network_blocks.py:31-213``,
  ``yolox/models/darknet.py:97-180``, ``yolox/utils/utils_snn.py:16-58``,
  ``yolox/utils/model_utils.py:35-77``; the conv/BN arithmetic is PINNED by golden
  vectors from the reference model run through ``oracle/sj_shim`` (so the neuron inside
  is the unpinned restatement above).
* ``oracle.detector`` follows ``yolox/models/spiking_yolo_pafpn.py:89-120``,
  ``yolox/models/yolo_head.py:141-250``, ``yolox/models/spiking_yolox.py:38-74``          -- PINNED
  by ``tests/golden/detector.npz``: the reference's own ``EventExp.get_model()`` (use_spike True,
  tiny width) run through ``oracle/sj_shim`` from histograms to decoded predictions and its
  ``postprocess`` detections.
* ``oracle.letterbox`` follows ``yolox/data/datasets/gen1.py:423-483`` (letterbox geometry + paste) and restates
  ``cv2.resize(INTER_LINEAR)`` (third-party opencv-python, present in the authoring container)      -- PINNED
  by ``tests/golden/letterbox.npz``: the reference method (and so cv2) run on synthetic frames; identical in float64.
"""
