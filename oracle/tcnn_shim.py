"""ORACLE support (test infrastructure, NOT product code) - a `tinycudann`-API
compatible shim in plain PyTorch.

Purpose: tiny-cuda-nn is an un-vendored, un-pinned external CUDA package
(/root/reference/README.md:88-91) that exists neither in the build container
nor on the GPU box.  Installing this module as ``sys.modules['tinycudann']``
lets the reference's own model/lidar4d.py, hash_field.py and flow_field.py be
imported and executed UNCHANGED on CPU, which is how tests/golden/make_golden.py
pins oracle/lidar4d_oracle.py against the reference code.

API surface used by the reference:
  tcnn.Encoding(n_input_dims, encoding_config)       hash_field.py:47,107; flow_field.py:67; lidar4d.py:68
  tcnn.Network(n_input_dims, n_output_dims, network_config)   lidar4d.py:83,95,107
  attributes .n_output_dims, flat fp32 nn.Parameter .params
The arithmetic is the spec in oracle/lidar4d_oracle.py (fp16-rounded tables,
fp32 blends and MLPs); outputs are fp32 (tcnn emits fp16).
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn

from lidar4d_b200.geometry import make_grid
from oracle import lidar4d_oracle as O


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, dtype=None, seed=1337):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.cfg = dict(encoding_config)
        otype = self.cfg["otype"]
        if otype == "HashGrid":
            self.geo = make_grid(
                n_input_dims, int(self.cfg["n_levels"]), int(self.cfg["n_features_per_level"]),
                int(self.cfg["log2_hashmap_size"]), int(self.cfg["base_resolution"]),
                float(self.cfg["per_level_scale"]))
            self.n_output_dims = self.geo.n_output_dims
            self.params = nn.Parameter(torch.empty(self.geo.n_params).uniform_(-1e-4, 1e-4))
        elif otype == "Frequency":
            self.degree = int(self.cfg.get("degree", self.cfg.get("n_frequencies", 12)))
            self.n_output_dims = n_input_dims * 2 * self.degree
            self.params = nn.Parameter(torch.zeros(0))
        else:
            raise NotImplementedError(otype)
        self.table_dtype = "fp16"

    def forward(self, x):
        if self.cfg["otype"] == "HashGrid":
            return O.hash_encode(x.float(), self.params, self.geo, self.table_dtype)
        return O.frequency_encode(x.float(), self.degree)


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        cfg = dict(network_config)
        assert cfg["otype"] == "FullyFusedMLP" and cfg["activation"] == "ReLU"
        assert cfg["output_activation"] == "None"
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.hidden = int(cfg["n_neurons"])
        self.n_hidden_layers = int(cfg["n_hidden_layers"])
        n_in_pad = (n_input_dims + 15) // 16 * 16
        assert n_output_dims <= 16
        self.params = nn.Parameter(O.OracleLiDAR4D._xavier_mlp(n_in_pad, self.hidden, self.n_hidden_layers))

    def forward(self, x):
        return O.fused_mlp(x.float(), self.params, self.n_input_dims, self.n_output_dims,
                           self.hidden, self.n_hidden_layers)


def install() -> types.ModuleType:
    """Register this shim as the `tinycudann` module."""
    mod = types.ModuleType("tinycudann")
    mod.Encoding = Encoding
    mod.Network = Network
    mod.__shim__ = True
    sys.modules["tinycudann"] = mod
    return mod
