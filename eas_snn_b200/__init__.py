"""eas_snn_b200 -- B200 (sm_100a) implementation of the EAS-SNN data-parallel hot path.

event binning -> adaptive spiking sampler -> multi-step LIF (+ conv) of the spiking backbone,
behind the reference's own module surface.  The kernels live in ``lib/libeas_b200.so`` (C ABI:
``include/eas_b200.h``); this package is the thin PyTorch host side.  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .binning import bin_events, rvt_event_sum, voxel_grid, letterbox_frames, HostEventBatch, CompactHist, poll_compact  # noqa: F401
from .embedding import AdaptiveRSNNEmbedding, SpikeCountEmbedding, LIFEmbedding, SpikingEmbedding  # noqa: F401
from .psee import DatRecording, bin_dat, dat_windows, pack_records, read_dat  # noqa: F401
from . import detector  # noqa: F401  (SpikingYOLOX / SpikingYOLOPAFPN / YOLOXHead / build_syolox / postprocess)
from .neuron import ATan, Sigmoid, Rect, ParametricLIFNode, plif_multistep, is_spiking_neuron, reset_net  # noqa: F401

__version__ = "0.1.0"
