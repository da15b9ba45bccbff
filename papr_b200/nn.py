"""Parameter containers that keep the reference's module tree, hence its state_dict keys.

Reference layout (models/attn.py, models/mlp.py):
  proximity_attn.embed.embed_{k,q,v}.{innorm,outnorm}.{a_2,b_2}
  proximity_attn.embed.embed_{k,q,v}.mlp.model.{1,3,5,...}.{weight,bias}      (odd entries are the Linear layers)
  proximity_attn.attention_layer.{w_k,w_q}.{weight,bias}
  mapping_mlp.model.model.{1,3,...}.{weight,bias}
The modules here own the parameters and small fp32 torch forwards (used for per-ray query work and the fp32
parity mode); the per-(ray,point) hot path reads the parameters from papr_b200.attention.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def make_activation(name):
    name = name.lower()
    if name == "none":
        return nn.Identity()
    if name == "relu":
        return nn.ReLU()
    if name == "leakyrelu":
        return nn.LeakyReLU(0.2)
    if name == "relu+1":
        return ReluPlusOne()
    if name == "sigmoid":
        return nn.Sigmoid()
    if name == "tanh":
        return nn.Tanh()
    if name == "clamp":
        return Clamp01()
    raise NotImplementedError(f"activation layer [{name}] is not supported by the B200 path")


def activation_slope(name):
    """Negative slope of a piecewise-linear activation, or None for identity."""
    name = name.lower()
    if name == "none":
        return None
    if name == "relu":
        return 0.0
    if name == "leakyrelu":
        return 0.2
    raise NotImplementedError(f"activation [{name}] is not supported inside the fused MLP stacks")


class ReluPlusOne(nn.Module):
    def forward(self, x):
        return F.relu(x) + 1.0


class Clamp01(nn.Module):
    def forward(self, x):
        return torch.clamp(x, 0, 1)


class LayerNorm(nn.Module):
    """attn.py:30-42: a_2 * (x - mean) / (std + eps) + b_2 with the unbiased std."""

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(features))
        self.b_2 = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        mean = x.mean(-1, keepdim=True)
        std = x.std(-1, keepdim=True)
        return self.a_2 * (x - mean) / (std + self.eps) + self.b_2


class MLP(nn.Module):
    """mlp.py:12-59 container: model = [Identity, Linear, act, Linear, act, ...]; xavier-uniform weights."""

    def __init__(self, inp_dim, num_layers, num_channels, out_dim, act_type="relu", last_act_type="none",
                 skip_layers=(), use_wn=False, half_layers=(), residual_layers=()):
        super().__init__()
        if use_wn or len(half_layers) or len(residual_layers):
            raise NotImplementedError("weight norm / half / residual layers are not used by any shipped config")
        self.inp_dim, self.out_dim, self.num_layers = inp_dim, out_dim, num_layers
        self.act_type, self.last_act_type = act_type, last_act_type
        self.skip_layers = list(skip_layers)
        layers = [nn.Identity()]
        for i in range(num_layers):
            cur_in = inp_dim if i == 0 else num_channels
            cur_out = out_dim if i == num_layers - 1 else num_channels
            if i in self.skip_layers:
                cur_in += inp_dim
            layers.append(nn.Linear(cur_in, cur_out))
            layers.append(make_activation(last_act_type if i == num_layers - 1 else act_type))
        self.model = nn.ModuleList(layers)
        for p in self.model.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def linears(self):
        return [self.model[2 * i + 1] for i in range(self.num_layers)]

    def forward(self, x):
        inp = x
        for i in range(self.num_layers):
            if i in self.skip_layers:
                x = torch.cat([x, inp], dim=-1)
            x = self.model[2 * i + 2](self.model[2 * i + 1](x))
        return x


class FeedForward(nn.Module):
    """attn.py:90-117 with dropout 0 / residual false (the only shipped setting)."""

    def __init__(self, d_input, opt, eps):
        super().__init__()
        if opt.residual_ff or opt.dropout_ff != 0.0:
            raise NotImplementedError("residual_ff / dropout are not used by any shipped config")
        self.d_input, self.d_output = d_input, opt.d_ff_out
        if opt.norm == "layernorm":
            self.innorm, self.outnorm = LayerNorm(d_input, eps), LayerNorm(opt.d_ff_out, eps)
        elif opt.norm == "none":
            self.innorm, self.outnorm = nn.Identity(), nn.Identity()
        else:
            raise ValueError("Invalid attention norm type")
        self.mlp = MLP(d_input, opt.n_ff_layer, opt.d_ff, opt.d_ff_out, opt.ff_act, opt.ff_last_act,
                       opt.skip_layers, opt.use_wn, opt.half_layers, opt.residual_layers)

    def forward(self, x):
        return self.outnorm(self.mlp(self.innorm(x)))


class Embeddings(nn.Module):
    def __init__(self, dk, dq, dv, embed_opt, eps):
        super().__init__()
        self.embed_k = FeedForward(dk, embed_opt.key, eps)
        self.embed_q = FeedForward(dq, embed_opt.query, eps)
        self.embed_v = FeedForward(dv, embed_opt.value, eps)


class AttentionLayer(nn.Module):
    def __init__(self, embed_opt, d_model, score_act):
        super().__init__()
        self.d_model = d_model
        self.w_k = nn.Linear(embed_opt.key.d_ff_out, d_model)
        self.w_q = nn.Linear(embed_opt.query.d_ff_out, d_model)
        nn.init.xavier_uniform_(self.w_k.weight)
        nn.init.xavier_uniform_(self.w_q.weight)
        self.score_act_type = score_act


class MappingMLP(nn.Module):
    """mlp.py:62-78 (shading code -> gamma | beta); runs once per step on a 128-vector, stays in PyTorch."""

    def __init__(self, opt, inp_dim, out_dim):
        super().__init__()
        self.model = MLP(inp_dim, opt.num_layers, opt.dim, out_dim, opt.act, opt.last_act, use_wn=opt.use_wn)

    def forward(self, x):
        return self.model(x)
