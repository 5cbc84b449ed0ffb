from .tcn import TCN, TCNBlock
from .gcn import GCN, GCNBlock
from .wavenet import WaveNet, WaveNet1dBlock, Conv1dStack
from .custom_layers import Conv1dCausal, FiLM, GatedAF, TanhAF

__all__ = ["TCN", "TCNBlock", "GCN", "GCNBlock", "WaveNet", "WaveNet1dBlock", "Conv1dStack", "Conv1dCausal", "FiLM", "GatedAF", "TanhAF"]
