from .tcn import TCN, TCNBlock
from .gcn import GCN, GCNBlock
from .custom_layers import Conv1dCausal, FiLM, GatedAF, TanhAF

__all__ = ["TCN", "TCNBlock", "GCN", "GCNBlock", "Conv1dCausal", "FiLM", "GatedAF", "TanhAF"]
