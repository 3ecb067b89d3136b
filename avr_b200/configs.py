"""The render / model dictionaries of the four reference configs BASELINE.json names.

Transcribed (values only) from ``/root/reference/config_files``:
``avr_simu.yml:11-21,44-88``, ``avr_meshrir.yml:11-21,45-89``,
``avr_raf_furnished.yml:11-22,42-104``, ``avr_real_exp_ch_emb_1.yml:11-21,45-93``.
The YAMLs themselves do not travel to the GPU box, so ``bench.py`` and the tests take the
shapes from here; ``tests/test_configs_match_reference.py`` diffs them against the YAMLs
whenever the reference tree is present.
"""
from __future__ import annotations

import copy


def _grid(log2_size: int = 18) -> dict:
    return {"base_resolution": 16, "log2_hashmap_size": log2_size, "n_features_per_level": 2,
            "n_levels": 20, "otype": "HashGrid"}


def _mlp(width: int, hidden: int, otype: str) -> dict:
    return {"activation": "ReLU", "n_hidden_layers": hidden, "n_neurons": width, "otype": otype,
            "output_activation": "None"}


def _avr_model(T: int, dir_log2: int = 18) -> dict:
    return {
        "signal_output_dim": T, "leaky_relu": 0.03,
        "pos_encoding_sigma": _grid(), "dir_encoding_sig": _grid(dir_log2), "tx_encoding_sig": _grid(),
        "sigma_encoder_network": _mlp(128, 3, "FullyFusedMLP"),
        "sigma_decoder_network": _mlp(128, 3, "FullyFusedMLP"),
        "signal_network": _mlp(512, 3, "CutlassMLP"),
    }


CONFIGS = {
    "simu": {
        "dataset_type": "Simu", "model_class": "AVRModel",
        "render": {"xyz_min": -10, "xyz_max": 10, "near": 0, "far": 6, "n_samples": 64, "n_azi": 64,
                   "n_ele": 32, "speed": 343.8, "fs": 16000, "pathloss": 1.5},
        "model": _avr_model(1600),
    },
    "meshrir": {
        "dataset_type": "MeshRIR", "model_class": "AVRModel",
        "render": {"xyz_min": -6, "xyz_max": 6, "near": 0, "far": 4, "n_samples": 64, "n_azi": 80,
                   "n_ele": 40, "speed": 343.8, "fs": 24000, "pathloss": 1.5},
        "model": _avr_model(2400, dir_log2=20),
    },
    "raf_furnished": {
        "dataset_type": "RAF", "model_class": "AVRModel_complex",
        "render": {"xyz_min": -12, "xyz_max": 12, "near": 0, "far": 6, "n_samples": 32, "n_azi": 36,
                   "n_ele": 18, "speed": 346.8, "fs": 16000, "pathloss": 0.5, "sig_length": 1600},
        "model": {
            "signal_output_dim": 1600, "leaky_relu": 0.03,
            "pos_encoding_sigma": _grid(), "pos_encoding_sig": _grid(), "dir_encoding_sig": _grid(),
            "tx_pos_encoding_sigma": _grid(), "tx_pos_encoding_sig": _grid(), "tx_dir_encoding_sig": _grid(),
            "sigma_encoder_network": _mlp(128, 3, "FullyFusedMLP"),
            "sigma_decoder_network": _mlp(128, 1, "FullyFusedMLP"),
            "signal_network": _mlp(512, 4, "CutlassMLP"),
        },
    },
    "real_exp_ch_emb_1": {
        "dataset_type": "Real_env", "model_class": "AVRModel",
        "render": {"xyz_min": 0, "xyz_max": 10, "near": 0, "far": 6, "n_samples": 64, "n_azi": 64,
                   "n_ele": 32, "speed": 343.8, "fs": 16000, "pathloss": 1.5},
        "model": dict(_avr_model(1600), channel_embed={"is_embed": True, "ch_num": 8, "emb_dim": 128}),
    },
}


def get_config(name: str) -> dict:
    return copy.deepcopy(CONFIGS[name])


def tiny_config(model_class: str = "AVRModel", *, n_azi=6, n_ele=3, n_samples=8, T=200, n_levels=6,
                log2_hashmap_size=10, base_resolution=4, width_sigma=32, width_signal=64,
                xyz_min=-10, xyz_max=10, far=6, fs=16000, speed=343.8, pathloss=1.5) -> dict:
    """A seconds-on-CPU shape with the same structure (dense + hashed levels, padding rules)."""
    grid = {"base_resolution": base_resolution, "log2_hashmap_size": log2_hashmap_size,
            "n_features_per_level": 2, "n_levels": n_levels, "otype": "HashGrid"}
    render = {"xyz_min": xyz_min, "xyz_max": xyz_max, "near": 0, "far": far, "n_samples": n_samples,
              "n_azi": n_azi, "n_ele": n_ele, "speed": speed, "fs": fs, "pathloss": pathloss}
    if model_class == "AVRModel":
        model = {"signal_output_dim": T, "leaky_relu": 0.03,
                 "pos_encoding_sigma": dict(grid), "dir_encoding_sig": dict(grid), "tx_encoding_sig": dict(grid),
                 "sigma_encoder_network": _mlp(width_sigma, 3, "FullyFusedMLP"),
                 "sigma_decoder_network": _mlp(width_sigma, 3, "FullyFusedMLP"),
                 "signal_network": _mlp(width_signal, 3, "CutlassMLP")}
    else:
        model = {"signal_output_dim": T, "leaky_relu": 0.03,
                 "pos_encoding_sigma": dict(grid), "pos_encoding_sig": dict(grid), "dir_encoding_sig": dict(grid),
                 "tx_pos_encoding_sigma": dict(grid), "tx_pos_encoding_sig": dict(grid),
                 "tx_dir_encoding_sig": dict(grid),
                 "sigma_encoder_network": _mlp(width_sigma, 3, "FullyFusedMLP"),
                 "sigma_decoder_network": _mlp(width_sigma, 1, "FullyFusedMLP"),
                 "signal_network": _mlp(width_signal, 4, "CutlassMLP")}
    return {"dataset_type": "tiny", "model_class": model_class, "render": render, "model": model}
