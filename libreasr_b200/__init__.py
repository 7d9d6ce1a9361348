"""libreasr_b200 -- B200-native (sm_100a CUDA) streaming RNN-Transducer inference behind
LibreASR's module surface.  See DESIGN.md / INTEGRATION.md.

Public surface
--------------
``LibreASR``                      facade named by BASELINE.json: ``transcribe()`` / ``stream()``
``libreasr_b200.lib.models``      ``Transducer`` / ``Encoder`` / ``Predictor`` / ``Joint`` (reference names)
``libreasr_b200.lib.transforms``  the inference transforms (reference names)
``libreasr_b200.engine.Engine``   typed veneer over the C ABI (include/rnnt_b200.h)
``libreasr_b200.parallel``        utterance sharding over the GPUs of one box (NCCL scatter / gather)
"""
from .engine import Engine, EngineConfig, tokens_to_lists  # noqa: F401
from .api import LibreASR, StreamBatch  # noqa: F401

__version__ = "0.1.0"
