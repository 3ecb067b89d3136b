"""The neural acoustic field, B200-native: drop-in for the reference ``model.py``.

``Encoding`` / ``Network`` stand where the reference instantiates ``tcnn.Encoding`` / ``tcnn.Network``
(``/root/reference/model.py:21,43,66-68,117,146,176,258-285``): one flat fp32 ``params`` Parameter per
module, tcnn's layout, so state-dict keys (``_pos_encoding.params`` ...) line up with
the reference's checkpoints.  ``AVRModel`` / ``AVRModel_complex`` keep the reference constructor
(``cfg`` = the YAML ``model:`` section) and ``forward`` signatures (``model.py:183,291``).

Called on its own, ``forward(pts, view, tx, ...)`` evaluates the field on arbitrary points with the
hand-written kernels (hash-grid gather, fp32 GEMMs) under autograd.  When wrapped by
``avr_b200.AVRRender`` the renderer instead asks for ``fused_plan()`` and runs the whole
ray-generation -> encode -> MLP -> composite pipeline without ever materialising the per-point inputs.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


INDEX_STRIDES = ("uint32", "exact")


def hashgrid_geometry(cfg: dict) -> dict:
    """Level table of a tcnn ``HashGrid`` config (tcnn defaults for absent keys).

    ``index_stride`` (not a tcnn key; default ``"uint32"``) selects the arithmetic of the dense-index stride in
    ``grid_index``: ``"uint32"`` wraps like tiny-cuda-nn's own ``uint32_t stride`` -- levels with
    ``2^16 <= res <= hashmap size`` (12..14 of the 2^18 grids, 12..16 of a 2^20 grid) are then indexed
    ``(x + y*res) % size`` instead of being hashed, which is what a tcnn-trained checkpoint expects; ``"exact"``
    keeps the stride in 64 bits (those levels are hashed)."""
    if cfg.get("otype", "HashGrid") not in ("HashGrid", "Grid"):
        raise NotImplementedError(f"encoding otype {cfg.get('otype')!r} is not built")
    n_levels = int(cfg.get("n_levels", 16))
    n_feat = int(cfg.get("n_features_per_level", 2))
    log2_size = int(cfg.get("log2_hashmap_size", 19))
    base = int(cfg.get("base_resolution", 16))
    log2_pls = np.float32(math.log2(float(cfg.get("per_level_scale", 2.0))))
    scale, res, size, offset, off = [], [], [], [], 0
    for lvl in range(n_levels):
        s = np.float32(np.exp2(np.float32(lvl) * log2_pls, dtype=np.float32) * np.float32(base) - np.float32(1.0))
        r = int(math.ceil(float(s))) + 1
        n = min(_round_up(r ** 3, 8), 1 << log2_size)
        scale.append(float(s)); res.append(r); size.append(n); offset.append(off)
        off += n
    index_stride = cfg.get("index_stride", "uint32")
    if index_stride not in INDEX_STRIDES:
        raise ValueError(f"index_stride must be one of {INDEX_STRIDES}")
    return {"n_levels": n_levels, "n_feat": n_feat, "scale": scale, "res": res, "size": size, "offset": offset,
            "total": off, "index_stride": index_stride}


class Encoding(nn.Module):
    """Multiresolution hash grid on 3-D unit-cube inputs (``tcnn.Encoding(3, cfg)``)."""

    def __init__(self, n_input_dims: int, encoding_config: dict, dtype=torch.float32, seed: int = 1337,
                 index_stride: str = None):
        super().__init__()
        if n_input_dims != 3:
            raise NotImplementedError("only 3-D hash grids are built")
        if index_stride is not None:
            encoding_config = dict(encoding_config, index_stride=index_stride)
        self.geom = hashgrid_geometry(encoding_config)
        self.n_input_dims = 3
        self.n_output_dims = self.geom["n_levels"] * self.geom["n_feat"]
        g = torch.Generator().manual_seed(seed)
        n = self.geom["total"] * self.geom["n_feat"]
        self.params = nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * 1e-4)      # tcnn: U(-1e-4, 1e-4)
        self._meta = None
        #: how the table gradient is accumulated: "deterministic" (default; int64 fixed point, bit-reproducible, as the
        #: north_star asks) or "atomic" (fp32 vector reductions: faster, run-to-run differences in the last bits like
        #: tcnn's own atomics)
        self.grid_grad = "deterministic"

    @property
    def meta(self):
        if self._meta is None:
            from . import ops
            self._meta = ops.make_grid_meta(self.geom)
        return self._meta

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        return Fn.HashGridFunction.apply(u.contiguous().float(), self.params, self)


class Network(nn.Module):
    """Bias-free ReLU MLP (``tcnn.Network(n_in, n_out, cfg)``).

    ``n_hidden_layers = h`` gives ``h+1`` row-major ``[out, in]`` matrices; the input is padded with ones
    and the output with unused rows up to a multiple of 16 (FullyFusedMLP) or 8 (CutlassMLP).
    """

    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: dict, seed: int = 1337):
        super().__init__()
        act = network_config.get("activation", "ReLU")
        out_act = network_config.get("output_activation", "None")
        if act != "ReLU" or out_act not in ("None", None):
            raise NotImplementedError("only ReLU hidden / linear output MLPs are built")
        align = 16 if network_config.get("otype", "FullyFusedMLP") == "FullyFusedMLP" else 8
        self.width = int(network_config["n_neurons"])
        self.n_hidden = int(network_config["n_hidden_layers"])
        if self.n_hidden < 0:
            raise ValueError("n_hidden_layers must be >= 0")
        if self.width % 4:
            raise NotImplementedError("n_neurons must be a multiple of 4")
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.in_pad, self.out_pad = _round_up(n_input_dims, align), _round_up(n_output_dims, align)
        dims = [self.in_pad] + [self.width] * self.n_hidden + [self.out_pad]
        self.shapes = [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]
        g = torch.Generator().manual_seed(seed)
        chunks = [(torch.rand(o * i, generator=g) * 2 - 1) * math.sqrt(6.0 / (i + o)) for (o, i) in self.shapes]
        self.params = nn.Parameter(torch.cat(chunks))                                # Xavier-uniform

    def matrices(self, flat: torch.Tensor):
        out, off = [], 0
        for (o, i) in self.shapes:
            out.append(flat[off:off + o * i].view(o, i))
            off += o * i
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return Fn.MLPFunction.apply(x.contiguous().float(), self.params, self)


class InjectionNetwork(nn.Module):
    """``LayeredTCNNWithInjection`` (``/root/reference/model.py:11-61``): every hidden layer is its own single-matrix
    network (``n_hidden_layers: 0``, linear) followed by ``h + layer_embeddings[l][ch_id]`` and ReLU; a linear
    output layer.  Sub-module / parameter names follow the reference (``hidden_layers.{l}.params``,
    ``layer_embeddings.{l}``, ``output_layer.params``).

    For the fused render step the layers present themselves as one ``Network``-shaped stack (``shapes``,
    ``matrices``, a concatenated ``params``); the embedding rows enter the tensor-core GEMM epilogue as a
    per-receiver bias (``AVR_UMMA_BIAS``).
    """

    def __init__(self, n_input_dims: int, n_neurons: int, n_hidden_layers: int, n_output_dims: int, ch_num: int,
                 activation: str = "ReLU", otype: str = "FullyFusedMLP", seed: int = 1337):
        super().__init__()
        if activation != "ReLU":
            raise NotImplementedError("only ReLU")
        if n_hidden_layers < 1:
            raise ValueError("n_hidden_layers must be >= 1")
        one = {"otype": otype, "activation": "ReLU", "output_activation": "None", "n_neurons": n_neurons,
               "n_hidden_layers": 0}
        align = 16 if otype == "FullyFusedMLP" else 8
        if n_neurons % align:
            raise NotImplementedError("n_neurons must be a multiple of the tcnn padding (16 / 8)")
        self.hidden_layers = nn.ModuleList()
        self.layer_embeddings = nn.ParameterList()
        g = torch.Generator().manual_seed(seed + 100)
        in_dim = n_input_dims
        for i in range(n_hidden_layers):
            self.hidden_layers.append(Network(in_dim, n_neurons, one, seed + 10 * (i + 1)))
            self.layer_embeddings.append(nn.Parameter(torch.randn(ch_num, n_neurons, generator=g) / math.sqrt(n_neurons)))
            in_dim = n_neurons
        self.output_layer = Network(in_dim, n_output_dims, one, seed + 10 * (n_hidden_layers + 1))
        self.width, self.n_hidden = int(n_neurons), int(n_hidden_layers)
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.in_pad, self.out_pad = self.hidden_layers[0].in_pad, self.output_layer.out_pad
        self.shapes = [layer.shapes[0] for layer in self._layers()]

    def _layers(self):
        return list(self.hidden_layers) + [self.output_layer]

    @property
    def params(self) -> torch.Tensor:
        """All matrices as one flat tensor (autograd splits the gradient back to the per-layer parameters)."""
        return torch.cat([layer.params for layer in self._layers()])

    matrices = Network.matrices

    def bias_rows(self, ch_idx: torch.Tensor):
        """Per-receiver pre-activation bias of every hidden layer: ``[bs, n_neurons]`` each."""
        return [emb[ch_idx] for emb in self.layer_embeddings]

    def forward(self, x: torch.Tensor, ch_id=None) -> torch.Tensor:
        for idx, layer in enumerate(self.hidden_layers):
            h = layer(x)
            if ch_id is not None:
                h = h + self.layer_embeddings[idx][ch_id]
            x = torch.relu(h)
        return self.output_layer(x)


class RowsInput:
    """A per-receiver ``[bs, width]`` block of a network input (a channel-embedding row, 'concat' mode)."""

    def __init__(self, key: str, width: int):
        self.key, self.n_output_dims = key, int(width)


def channel_embed_modes(cfg: dict):
    """``model.py:71-90`` -> ({enc,dec,sig: 'injection'|'concat'|'none'}, embedding widths, ch_num)."""
    ch = cfg.get("channel_embed") or {}
    is_embed, conn = ch.get("is_embed", False), ch.get("connection_type", None)
    mode = {}
    for k, name in (("enc", "sigma_encoder"), ("dec", "sigma_decoder"), ("sig", "signal_network")):
        on = is_embed and bool(ch.get("is_" + name, False))
        mode[k] = "injection" if on and conn == "add" else "concat" if on and conn == "concat" else "none"
    dims = {"enc": int(ch.get("emb_dim_sigma_encoder", 0)), "dec": int(ch.get("emb_dim_sigma_decoder", 0)),
            "sig": int(ch.get("emb_dim_signal_network", 0))}
    return mode, dims, int(ch.get("ch_num", 0))


class AVRModel(nn.Module):
    """``/root/reference/model.py:63-235`` (MeshRIR / Simu / Real_env), channel-embedding variants included
    (``channel_embed.connection_type: add | concat``, ``model.py:71-181,193-228``)."""

    _NETS = {"enc": ("_model_encoder_sigma", "encoder_channel_embedding", "sigma_encoder_network", "FullyFusedMLP"),
             "dec": ("_model_decoder_sigma", "decoder_channel_embedding", "sigma_decoder_network", "FullyFusedMLP"),
             "sig": ("_model_signal", "signal_channel_embedding", "signal_network", "CutlassMLP")}

    def __init__(self, cfg: dict, seed: int = 1337):
        super().__init__()
        self._pos_encoding = Encoding(3, cfg["pos_encoding_sigma"], seed=seed)
        self._dir_encoding = Encoding(3, cfg["dir_encoding_sig"], seed=seed + 1)
        self._tx_encoding = Encoding(3, cfg["tx_encoding_sig"], seed=seed + 2)
        self.signal_output_dim = int(cfg["signal_output_dim"])
        self.embed_mode, emb_dims, self.ch_num = channel_embed_modes(cfg)
        self.encoder_mode, self.decoder_mode, self.signal_mode = (self.embed_mode[k] for k in ("enc", "dec", "sig"))
        base_in = {"enc": self._pos_encoding.n_output_dims, "dec": 128,
                   "sig": 128 + self._dir_encoding.n_output_dims + self._tx_encoding.n_output_dims}
        n_out = {"enc": 128, "dec": 1, "sig": self.signal_output_dim}
        g = torch.Generator().manual_seed(seed + 50)
        for k, off in (("enc", 3), ("dec", 4), ("sig", 5)):
            attr, emb_attr, cfg_key, default_otype = self._NETS[k]
            ncfg = cfg[cfg_key]
            if self.embed_mode[k] == "injection":
                net = InjectionNetwork(base_in[k], int(ncfg["n_neurons"]), int(ncfg["n_hidden_layers"]), n_out[k],
                                       self.ch_num, ncfg.get("activation", "ReLU"), ncfg.get("otype", default_otype),
                                       seed + off)
            else:
                n_in = base_in[k]
                if self.embed_mode[k] == "concat":
                    if emb_dims[k] < 1 or self.ch_num < 1:
                        raise ValueError("channel_embed concat needs ch_num and a positive emb_dim_* for the network")
                    setattr(self, emb_attr, nn.Parameter(torch.randn(self.ch_num, emb_dims[k], generator=g) /
                                                         math.sqrt(emb_dims[k])))
                    n_in += emb_dims[k]
                net = Network(n_in, n_out[k], ncfg, seed + off)
            setattr(self, attr, net)
        self.leaky_slope = 0.01      # model.py:233 -- F.leaky_relu default, cfg["leaky_relu"] is ignored

    def _run(self, k, x, ch):
        attr, emb_attr = self._NETS[k][:2]
        net = getattr(self, attr)
        if self.embed_mode[k] == "injection":
            return net(x, ch)
        if self.embed_mode[k] == "concat":
            if ch is None:
                raise ValueError("this field concatenates a channel embedding: ch_idx is required")
            x = torch.cat([x, getattr(self, emb_attr)[ch]], -1)
        return net(x)

    def forward(self, pts, view, tx, ch_idx=None):
        bs, n_pts = pts.size(0), pts.size(1)
        ch = ch_idx.unsqueeze(1).expand(-1, n_pts).reshape(-1) if ch_idx is not None else None
        u_pts = Fn.unit_cube(pts)
        sigma_feat = self._run("enc", self._pos_encoding(u_pts), ch)
        attn = self._run("dec", torch.relu(sigma_feat), ch)
        sig_in = torch.cat([sigma_feat, self._dir_encoding(Fn.unit_cube(view)), self._tx_encoding(Fn.unit_cube(tx))], -1)
        signal = self._run("sig", sig_in, ch)
        attn = torch.abs(torch.nn.functional.leaky_relu(attn, self.leaky_slope)).view(bs, n_pts, 1)
        return attn, signal.view(bs, n_pts, self.signal_output_dim)

    def small_table_modules(self):
        """Encodings evaluated on R / bs points only (per ray, per receiver): their table gradients are a few rows, which
        data-parallel runs all-gather instead of all-reducing the tables (``ddp.GradArena.attach``)."""
        return [self._dir_encoding, self._tx_encoding]

    def fused_plan(self, ch_idx=None) -> dict:
        """What the fused render step needs.  With a channel embedding the per-receiver rows (gathered here, under
        autograd) ride along as ``extra_tensors``: ``('bias', net, layer)`` rows are added to a hidden layer's
        pre-activation in the GEMM epilogue, ``('rows', key)`` blocks are broadcast into a network input."""
        plan = {
            "x0": [(self._pos_encoding, "point")], "dec_tail": [],
            "tail": [(self._dir_encoding, "ray"), (self._tx_encoding, "receiver_tx")],
            "enc": self._model_encoder_sigma, "dec": self._model_decoder_sigma, "sig": self._model_signal,
            "feat_dim": 128, "sig_relu_feat": False, "slope": self.leaky_slope, "needs_dir_tx": False,
            "extras": [], "extra_tensors": [],
        }
        for k, seg in (("enc", "x0"), ("dec", "dec_tail"), ("sig", "tail")):
            attr, emb_attr = self._NETS[k][:2]
            if self.embed_mode[k] == "injection" and ch_idx is not None:
                for li, rows in enumerate(getattr(self, attr).bias_rows(ch_idx)):
                    plan["extras"].append(("bias", k, li))
                    plan["extra_tensors"].append(rows)
            elif self.embed_mode[k] == "concat":
                if ch_idx is None:
                    raise ValueError("this field concatenates a channel embedding: ch_idx is required")
                emb = getattr(self, emb_attr)
                plan[seg].append((RowsInput(emb_attr, emb.shape[1]), "receiver_rows"))
                plan["extras"].append(("rows", emb_attr))
                plan["extra_tensors"].append(emb[ch_idx])
        return plan


class AVRModel_complex(nn.Module):
    """``/root/reference/model.py:238-331`` (RAF).  ``ch_idx`` is accepted and ignored."""

    def __init__(self, cfg: dict, seed: int = 1337):
        super().__init__()
        self.leaky_slope = float(cfg["leaky_relu"])
        self.signal_output_dim = int(cfg["signal_output_dim"])
        self._pos_encoding = Encoding(3, cfg["pos_encoding_sigma"], seed=seed)
        self._pos_signal_encoding = Encoding(3, cfg["pos_encoding_sig"], seed=seed + 1)
        self._tx_pos_encoding = Encoding(3, cfg["tx_pos_encoding_sigma"], seed=seed + 2)
        self._tx_pos_signal_encoding = Encoding(3, cfg["tx_pos_encoding_sig"], seed=seed + 3)
        self._dir_encoding = Encoding(3, cfg["dir_encoding_sig"], seed=seed + 4)
        self._tx_dir_encoding = Encoding(3, cfg["tx_dir_encoding_sig"], seed=seed + 5)
        n_enc = self._pos_encoding.n_output_dims
        self._model_encoder_sigma = Network(n_enc + self._tx_pos_encoding.n_output_dims, 256,
                                            cfg["sigma_encoder_network"], seed + 6)
        self._model_decoder_sigma = Network(256, 1, cfg["sigma_decoder_network"], seed + 7)
        n_sig = (256 + self._dir_encoding.n_output_dims + self._tx_dir_encoding.n_output_dims +
                 self._pos_signal_encoding.n_output_dims + self._tx_pos_signal_encoding.n_output_dims)
        self._model_signal = Network(n_sig, self.signal_output_dim, cfg["signal_network"], seed + 8)

    def forward(self, pts, view, tx, tx_view, ch_idx=None):
        bs, n_pts = pts.size(0), pts.size(1)
        u_pts, u_tx = Fn.unit_cube(pts), Fn.unit_cube(tx)
        sigma_feat = self._model_encoder_sigma(torch.cat([self._pos_encoding(u_pts), self._tx_pos_encoding(u_tx)], -1))
        attn = self._model_decoder_sigma(torch.relu(sigma_feat))
        feat = torch.cat([torch.relu(sigma_feat), self._dir_encoding(Fn.unit_cube(view)),
                          self._tx_dir_encoding(Fn.unit_cube(tx_view)), self._pos_signal_encoding(u_pts),
                          self._tx_pos_signal_encoding(u_tx)], -1)
        signal = self._model_signal(feat)
        attn = torch.abs(torch.nn.functional.leaky_relu(attn, self.leaky_slope)).view(bs, n_pts, 1)
        return attn, signal.reshape(bs, n_pts, self.signal_output_dim)

    def small_table_modules(self):
        return [self._tx_pos_encoding, self._dir_encoding, self._tx_dir_encoding, self._tx_pos_signal_encoding]

    def fused_plan(self) -> dict:
        return {
            "x0": [(self._pos_encoding, "point"), (self._tx_pos_encoding, "receiver_tx")],
            "tail": [(self._dir_encoding, "ray"), (self._tx_dir_encoding, "receiver_dir_tx"),
                     (self._pos_signal_encoding, "point"), (self._tx_pos_signal_encoding, "receiver_tx")],
            "enc": self._model_encoder_sigma, "dec": self._model_decoder_sigma, "sig": self._model_signal,
            "feat_dim": 256, "sig_relu_feat": True, "slope": self.leaky_slope, "needs_dir_tx": True,
        }
