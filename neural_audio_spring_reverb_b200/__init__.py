"""B200-native TCN / GCN spring-reverb inference engine.

Drop-in for the inference path of francescopapaleo/neural-audio-spring-reverb
(TCN/GCN.forward, load_model_checkpoint, make_inference, measure_rtf); the
arithmetic runs in hand-written sm_100a kernels behind the C ABI of
include/nasr_b200.h (libnasr_b200.so)."""
from .networks import TCN, GCN, WaveNet, TCNBlock, GCNBlock
from .networks.model_utils import initialize_model, load_model_checkpoint, parse_config
from .streaming import CachedStream

__version__ = "0.1.0"
__all__ = ["TCN", "GCN", "WaveNet", "TCNBlock", "GCNBlock", "initialize_model", "load_model_checkpoint",
           "parse_config", "CachedStream", "make_inference", "measure_rtf"]


def __getattr__(name):
    if name == "make_inference":
        from .inference import make_inference
        return make_inference
    if name == "measure_rtf":
        from .rtf import measure_rtf
        return measure_rtf
    raise AttributeError(name)
