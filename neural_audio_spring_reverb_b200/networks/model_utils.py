"""Model registry and checkpoint loader with the reference's signatures
(src/neural_audio_spring_reverb/networks/model_utils.py:13-165).

Only the two architectures on the accelerated path are constructible.  The
checkpoint format is the reference's torch.save dict (model_utils.py:206-216);
the writer (save_model_checkpoint) is training-side and out of scope."""
import torch
import yaml

from .gcn import GCN
from .tcn import TCN
from .wavenet import WaveNet

# constructor keyword allow-lists (model_utils.py:50-99)
_CTOR_KEYS = {
    "TCN": (TCN, {"n_channels", "n_layers", "dilation_growth", "in_ch", "out_ch", "kernel_size", "cond_dim"}),
    "GCN": (GCN, {"in_ch", "out_ch", "n_blocks", "n_channels", "dilation_growth", "kernel_size", "cond_dim"}),
    "WaveNet": (WaveNet, {"in_ch", "out_ch", "n_blocks", "n_stacks", "n_channels", "kernel_size", "dilation_growth",
                          "cond_dim"}),
}
_OUT_OF_SCOPE = {"LSTM", "GRU"}


def parse_config(config_path):
    """YAML file -> dict (model_utils.py:13-19)."""
    with open(config_path, "r") as fh:
        return dict(yaml.safe_load(fh))


def initialize_model(device, config):
    """Build the model named by config["model_type"]; returns (model, rf, params)
    like model_utils.py:22-124."""
    kind = config["model_type"]
    if kind in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f"model_type {kind!r} is not on the B200 inference path (TCN, GCN and WaveNet only)")
    if kind not in _CTOR_KEYS:
        raise ValueError(f"Unknown model type: {kind}")
    cls, allowed = _CTOR_KEYS[kind]
    kwargs = {k: v for k, v in config.items() if k in allowed}
    model = cls(**kwargs).to(device)
    print(f"Configuration name: {config['name']}")
    rf = model.calc_receptive_field()
    print(f"Receptive field: {rf} samples or {(rf / config['sample_rate'])*1e3:0.1f} ms")
    params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    print(f"Parameters: {params*1e-3:0.3f} k")
    return model, rf, params


def load_model_checkpoint(args):
    """args.checkpoint, args.device -> (model, optimizer_sd, scheduler_sd, config, rf, params)
    (model_utils.py:127-165). load_state_dict is strict."""
    checkpoint = torch.load(args.checkpoint, map_location=args.device)
    model_state_dict = checkpoint.get("model_state_dict")
    optimizer_state_dict = checkpoint.get("optimizer_state_dict", None)
    scheduler_state_dict = checkpoint.get("scheduler_state_dict", None)
    loaded_config = checkpoint["config_state_dict"]
    model, rf, params = initialize_model(args.device, loaded_config)
    model.load_state_dict(model_state_dict)
    return model, optimizer_state_dict, scheduler_state_dict, loaded_config, rf, params
