"""``CatanPolicy`` — the reference's policy / value network (RL/models/*, 1 928 995 parameters), restated for BATCHED
execution on the GPU (BASELINE configs 3-5).

Same architecture, same parameter names (the reference's ``state_dict`` — e.g. ``RL/results/default_after_update_3825.pt`` —
loads with ``load_reference_state_dict``), same outputs for the same inputs (values, joint log-probs and entropies within
1e-5 of ``SettlersAgentPolicy.evaluate_actions``: ``tests/test_policy_net_vs_reference.py``).  It stays PyTorch, as
``north_star`` asks; what changes is how it runs:

* the reference walks the batch in Python (one ``for b in range(B)`` loop and one ``.cpu()`` sync per card list and call,
  RL/models/player_modules.py:49-66) and was only ever run at batch 1 inside the workers (game_manager.py:85); here every
  step is a batched tensor op without host synchronisation, so a whole tick (``catan_policy_inputs`` -> ``act`` ->
  ``catan_step`` -> ``catan_rollout_store``) can be captured in ONE CUDA graph;
* the first layer of all twelve action heads reads the same 512-wide trunk output: their trunk columns are ONE
  [B, 512] x [512, 1536] GEMM; the few autoregressive inputs (2-12 columns: one-hots of earlier heads) are added as small
  rank-k terms instead of concatenating a new [B, 512 + k] input per head (action_heads_module.py:49-62);
* the four "played cards" attention calls (current player + three opponents share ``played_card_mha``) run as one batch;
* the tail of a head (mask, softmax, sample, log-prob, entropy: RL/distributions.py:11-40, ~10 launches) is the fused
  ``catan_masked_categorical`` kernel when sampling on CUDA; ``evaluate_actions`` keeps torch ops for autograd.

Inputs are what ``PolicyInputs`` produces from packed env rows: ``obs`` dict (numeric keys ``[B, ...]`` float, five card
lists int64 ``[B, 25]`` zero-padded) and ``masks`` = list of 12 with heads 1 / 6 / 9 as ``[types, B, dim]``
(policy.py:185-190).  Actions are the int32 ``[B, 20]`` rows ``VecCatanEnv.step`` takes.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import layout as L

TRUNK = 512
HEAD_OUT = (13, 54, 73, 19, 5, 2, 3, 6, 6, 5, 5, 5)          # build_agent_model.py:58-78
HEAD_EXTRA = (0, 2, 0, 0, 0, 32, 2, 6, 12, 4, 9, 0)          # columns of mlp_1 beyond the 512 trunk columns
PLACE_SETTLEMENT, PLACE_ROAD, UPGRADE_CITY, BUY_DEV, PLAY_DEV, EXCHANGE, PROPOSE, RESPOND, MOVE_ROBBER, ROLL, END_TURN, STEAL, DISCARD = range(13)
CARD_YOP, CARD_MONOPOLY = 2, 4


def _ortho(m: nn.Linear, gain: float = math.sqrt(2)) -> nn.Linear:
    nn.init.orthogonal_(m.weight.data, gain=gain)
    nn.init.constant_(m.bias.data, 0)
    return m


def _layer_norm_small(x: torch.Tensor, norm: nn.LayerNorm) -> torch.Tensor:
    """LayerNorm over a short last dimension (16 / 25 / 64) of a tensor with very many rows.  ``F.layer_norm`` launches one block
    per row, which at a few dozen elements per row runs at a few percent of the memory bandwidth (profiles/r2_notes.md: 31-40 % of
    a rollout tick); moments by a vectorised reduction + elementwise ops are several times faster and the same function."""
    if x.is_cuda:                                           # one fused launch (csrc/policy_kernels.cu), with its own backward
        from .policy_ops import layer_norm_small
        return layer_norm_small(x, norm.weight, norm.bias, norm.eps)
    x = x.float()
    var, mean = torch.var_mean(x, dim=-1, unbiased=False, keepdim=True)
    return (x - mean) * torch.rsqrt(var + norm.eps) * norm.weight.float() + norm.bias.float()


class _MHA(nn.Module):
    """multi_headed_attention.py:11-54; ``key_mask`` [B, L] bool = keys that may be attended"""

    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.heads, self.hd = heads, dim // heads
        self.qkv_nets = nn.ModuleList([nn.Linear(dim, dim) for _ in range(3)])
        self.out_proj_net = nn.Linear(dim, dim)

    def forward(self, x: torch.Tensor, key_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, S, D = x.shape
        w = torch.cat([n.weight for n in self.qkv_nets], dim=0)
        b = torch.cat([n.bias for n in self.qkv_nets], dim=0)
        qkv = F.linear(x, w, b)                                                                 # one GEMM for q, k, v
        if x.is_cuda and key_mask is None and (S, D, self.heads) == (19, 64, 4):
            from .policy_ops import tile_attention              # 19 tokens x 4 heads of 16: the library kernels tile 64 x 64
            return self.out_proj_net(tile_attention(qkv))
        qkv = qkv.view(B, S, 3, self.heads, self.hd).permute(2, 0, 3, 1, 4)
        am = None if key_mask is None else key_mask.view(B, 1, 1, S)
        y = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=am)
        return self.out_proj_net(y.transpose(1, 2).reshape(B, S, D))


class _FFN(nn.Module):
    def __init__(self, dim: int, mult: int):
        super().__init__()
        self.linear1, self.linear2 = _ortho(nn.Linear(dim, mult * dim)), _ortho(nn.Linear(mult * dim, dim))

    def forward(self, x):
        return self.linear2(F.relu(self.linear1(x)))


class _SubLayer(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.norm = nn.LayerNorm(dim)


class _EncoderLayer(nn.Module):
    """tile_encoder.py:43-57: pre-norm residual attention + pointwise net"""

    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.sublayers = nn.ModuleList([_SubLayer(dim), _SubLayer(dim)])
        self.multi_headed_attention = _MHA(dim, heads)
        self.pointwise_net = _FFN(dim, 2)

    def forward(self, x):
        x = x + self.multi_headed_attention(_layer_norm_small(x, self.sublayers[0].norm))
        return x + self.pointwise_net(_layer_norm_small(x, self.sublayers[1].norm))


class _TileEncoder(nn.Module):
    """tile_encoder.py:60-91"""

    def __init__(self, in_dim=60, dim=64, heads=4, layers=2, out_dim=25):
        super().__init__()
        self.first_layer = _ortho(nn.Linear(in_dim, dim))
        self.encoder_layers = nn.ModuleList([_EncoderLayer(dim, heads) for _ in range(layers)])
        self.norm, self.norm_2 = nn.LayerNorm(out_dim), nn.LayerNorm(dim)
        self.out_proj = _ortho(nn.Linear(dim, out_dim))

    def forward(self, tiles):                       # [B, 19, 60]
        x = F.relu(_layer_norm_small(self.first_layer(tiles), self.norm_2))
        for layer in self.encoder_layers:
            x = layer(x)
        return F.relu(_layer_norm_small(self.out_proj(x), self.norm).reshape(tiles.shape[0], -1))


class _CurrentPlayer(nn.Module):
    """player_modules.py:12-98 (parameters only; the forward is batched in ``_Observation``)"""

    def __init__(self, in_dim=152, card_dim=16, proj=25):
        super().__init__()
        self.main_input_layer_1 = _ortho(nn.Linear(in_dim, 256))
        self.norm, self.norm_1 = nn.LayerNorm(card_dim), nn.LayerNorm(256)
        self.norm_2, self.norm_3, self.norm_4 = nn.LayerNorm(proj), nn.LayerNorm(proj), nn.LayerNorm(128)
        self.proj_hidden_dev_card, self.proj_played_dev_card = _ortho(nn.Linear(card_dim, proj)), _ortho(nn.Linear(card_dim, proj))
        self.final_linear_layer = _ortho(nn.Linear(2 * proj + 256, 128))


class _OtherPlayers(nn.Module):
    """player_modules.py:101-157"""

    def __init__(self, in_dim=159, card_dim=16, proj=25):
        super().__init__()
        self.main_input_layer_1 = _ortho(nn.Linear(in_dim, 256))
        self.proj_played_dev_card = _ortho(nn.Linear(card_dim, proj))
        self.final_linear_layer = _ortho(nn.Linear(proj + 256, 128))
        self.norm, self.norm_1, self.norm_2, self.norm_3 = nn.LayerNorm(card_dim), nn.LayerNorm(256), nn.LayerNorm(proj), nn.LayerNorm(128)


class _Observation(nn.Module):
    """observation_module.py:16-61"""

    def __init__(self):
        super().__init__()
        self.dev_card_embedding = nn.Embedding(6, 16)
        self.hidden_card_mha, self.played_card_mha = _MHA(16, 4), _MHA(16, 4)
        self.tile_encoder = _TileEncoder()
        self.current_player_module, self.other_players_module = _CurrentPlayer(), _OtherPlayers()
        self.final_layer = _ortho(nn.Linear(19 * 25 + 4 * 128, TRUNK))
        self.norm = nn.LayerNorm(TRUNK)

    def _card_lists(self, cards: torch.Tensor, mha: _MHA, norms) -> torch.Tensor:
        """``cards`` [G, B, 25] (G lists per sample, zero-padded card codes) -> [G, B, 16]: what the reference computes as embedding ->
        masked self-attention over the list -> LayerNorm (``norms[g]``) -> zero the padding -> sum over the positions
        (player_modules.py:49-84), WITHOUT materialising the sequence.  There is no positional term, so the attention output of a
        position depends only on its own card kind and on how many cards of each kind the list holds: with counts c[k] of the six
        token kinds (an empty list is one padding token, :53-55), the softmax weight of kind k for a query of kind a is
        c[k] exp(s[a, k]) / sum_k' c[k'] exp(s[a, k']) with the 6 x 6 score table s of the embedding, and the pooled output is
        sum_a c[a] LayerNorm(out_proj(attention[a])).  Two GEMMs with K = 6 replace the [B, 25, 16] sequence tensors; the result
        is the same function (sums re-associated)."""
        with torch.autocast(device_type=cards.device.type, enabled=False):      # a few K = 6 products: kept in fp32
            return self._card_lists_fp32(cards, mha, norms)

    def _card_lists_fp32(self, cards: torch.Tensor, mha: _MHA, norms) -> torch.Tensor:
        G, B, S = cards.shape
        kinds = torch.arange(1, 6, device=cards.device)
        c = (cards.unsqueeze(-1) == kinds).sum(dim=2).to(torch.float32)                         # [G, B, 5]
        c = torch.cat(((c.sum(-1, keepdim=True) == 0).to(torch.float32), c), dim=-1)              # kind 0: the single padding token
        H, hd = mha.heads, mha.hd
        e = self.dev_card_embedding.weight.float()                                                # [6, 16]
        q, k, v = (F.linear(e, n.weight.float(), n.bias.float()).view(6, H, hd).transpose(0, 1) for n in mha.qkv_nets)   # [H, 6, hd]
        sc = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(hd)                                   # [H, a, k]
        p = torch.exp(sc - sc.amax(dim=-1, keepdim=True))
        m_num = (p.unsqueeze(-1) * v.unsqueeze(1)).permute(2, 0, 1, 3).reshape(6, H * 6 * hd)     # [k, (h, a, d)]
        m_den = p.permute(2, 0, 1).reshape(6, H * 6)                                              # [k, (h, a)]
        cf = c.view(G * B, 6)
        att = (cf @ m_num).view(-1, H, 6, hd) / (cf @ m_den).view(-1, H, 6, 1)                    # [GB, H, a, hd]
        att = att.permute(0, 2, 1, 3).reshape(G, B, 6, H * hd)
        out = F.linear(att, mha.out_proj_net.weight.float(), mha.out_proj_net.bias.float())      # [G, B, 6, 16]
        pooled = []
        for g in range(G):
            pooled.append((_layer_norm_small(out[g], norms[g]) * c[g].unsqueeze(-1)).sum(dim=1))
        return torch.stack(pooled, dim=0)

    def forward(self, obs: Dict[str, torch.Tensor]) -> torch.Tensor:
        B = obs["current_player_main"].shape[0]
        cur, oth = self.current_player_module, self.other_players_module
        tiles = self.tile_encoder(obs["tile_representations"])
        # played cards of all four players through the shared attention in one batch; each module has its own LayerNorm
        played = torch.stack([obs["current_player_played_dev"], obs["next_player_played_dev"], obs["next_next_player_played_dev"],
                              obs["next_next_next_player_played_dev"]], dim=0)
        pooled = self._card_lists(played, self.played_card_mha, (cur.norm, oth.norm, oth.norm, oth.norm))
        cur_played, oth_played = pooled[0], pooled[1:]
        cur_hidden = self._card_lists(obs["current_player_hidden_dev"].unsqueeze(0), self.hidden_card_mha, (cur.norm,))[0]
        cur_hidden = F.relu(_layer_norm_small(cur.proj_hidden_dev_card(cur_hidden), cur.norm_2))
        cur_played = F.relu(_layer_norm_small(cur.proj_played_dev_card(cur_played), cur.norm_3))
        cur_main = F.relu(cur.norm_1(cur.main_input_layer_1(obs["current_player_main"])))
        cur_out = F.relu(cur.norm_4(cur.final_linear_layer(torch.cat((cur_main, cur_played, cur_hidden), dim=-1))))
        oth_main = torch.stack([obs["next_player_main"], obs["next_next_player_main"], obs["next_next_next_player_main"]], dim=0)
        oth_main = F.relu(oth.norm_1(oth.main_input_layer_1(oth_main)))
        oth_played = F.relu(_layer_norm_small(oth.proj_played_dev_card(oth_played), oth.norm_2))
        oth_out = F.relu(oth.norm_3(oth.final_linear_layer(torch.cat((oth_main, oth_played), dim=-1))))     # [3, B, 128]
        final = torch.cat((tiles, cur_out, oth_out[0], oth_out[1], oth_out[2]), dim=-1)
        return F.relu(self.norm(self.final_layer(final)))


class _Cat(nn.Module):
    """RL/distributions.py:25-40 ``Categorical``: holds ``linear``"""

    def __init__(self, n_in: int, n_out: int):
        super().__init__()
        self.linear = _ortho(nn.Linear(n_in, n_out), gain=0.01)


class _Head(nn.Module):
    """action_heads_module.py:182-228 / :232-264 (parameters)"""

    def __init__(self, in_dim: int, out_dim: int, custom_in: int = 0, custom_out: int = 0):
        super().__init__()
        if custom_in:
            self.custom_mlp, self.custom_norm = nn.Linear(custom_in, custom_out), nn.LayerNorm(custom_out)
        self.mlp_1, self.mlp_2, self.norm = nn.Linear(in_dim, 128), nn.Linear(128, 128), nn.LayerNorm(128)
        self.distribution = _Cat(128, out_dim)


class _Heads(nn.Module):
    def __init__(self):
        super().__init__()
        self.action_heads = nn.ModuleList(
            [_Head(TRUNK + HEAD_EXTRA[h], HEAD_OUT[h], custom_in=12 if h == 5 else 0, custom_out=32 if h == 5 else 0) for h in range(12)])


class _ValueNormaliser(nn.Module):
    """RL/models/utils.py:8-20: the constants the reference actually uses are the floats taken at construction (150, 150)"""

    def __init__(self, mean=150.0, std=150.0):
        super().__init__()
        self.mean, self.std = nn.Parameter(mean * torch.ones(1)), nn.Parameter(std * torch.ones(1))
        self.mean_np, self.std_np = float(mean), float(std)

    def normalise(self, v):
        return (v - self.mean_np) / (self.std_np + 1e-4)

    def denormalise(self, v):
        return self.mean_np + v * self.std_np


def _entropy(logp_all: torch.Tensor) -> torch.Tensor:
    """FixedCategorical.entropy (distributions.py:18-20) from log-probabilities: illegal entries contribute nothing"""
    p = logp_all.exp()
    return -(p * torch.where(p > 0, logp_all, torch.zeros_like(logp_all))).sum(-1)


class CatanPolicy(nn.Module):
    include_lstm = False
    lstm_size = 256
    use_value_normalisation = True

    def __init__(self):
        super().__init__()
        self.value_normaliser = _ValueNormaliser()
        self.observation_module = _Observation()
        self.action_head_module = _Heads()
        self.value_network_fc_1, self.value_network_fc_2, self.value_out = nn.Linear(TRUNK, 256), nn.Linear(256, 128), nn.Linear(128, 1)
        self.v_norm_1, self.v_norm_2 = nn.LayerNorm(256), nn.LayerNorm(128)
        self.register_buffer("_corner_row", torch.tensor([0, 2, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2]), persistent=False)   # build_agent_model.py:113-127
        self.register_buffer("_player_row", torch.tensor([2, 2, 2, 2, 2, 2, 0, 2, 2, 2, 2, 1, 2]), persistent=False)
        self.register_buffer("_exch_row_type", torch.tensor([1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1]), persistent=False)
        self.register_buffer("_exch_row_card", torch.tensor([1, 1, 3, 1, 2]), persistent=False)

    # ------------------------------------------------------------------ reference checkpoints
    def load_reference_state_dict(self, sd: dict) -> None:
        """a ``SettlersAgentPolicy.state_dict()``: same keys, minus the reference's empty ``dummy_param`` entries"""
        self.load_state_dict({k: v for k, v in sd.items() if not k.endswith("dummy_param")}, strict=True)

    def reference_state_dict(self) -> dict:
        """the other way round (for ``SettlersAgentPolicy.load_state_dict(..., strict=False)``)"""
        return {k: v.detach().clone() for k, v in self.state_dict().items()}

    # ------------------------------------------------------------------ trunk + value (policy.py:60-69)
    def base(self, obs: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        main = self.observation_module(obs)
        v = F.relu(self.v_norm_1(self.value_network_fc_1(main)))
        v = self.value_out(F.relu(self.v_norm_2(self.value_network_fc_2(v))))
        return v, main

    def get_value(self, obs, hidden_states=None, nonterminal_masks=None) -> torch.Tensor:
        return self.base(obs)[0]

    # ------------------------------------------------------------------ the twelve heads (action_heads_module.py:25-179)
    def _run_heads(self, main, obs, masks, actions: Optional[torch.Tensor], deterministic: bool, generator):
        """``actions``: int64 [B, 20] rows to evaluate, or None to sample.  Returns (action rows int64 [B, 20], joint log-prob
        [B, 1], entropy scalar)."""
        B, dev = main.shape[0], main.device
        heads = self.action_head_module.action_heads
        sampling = actions is None
        fused = sampling and main.is_cuda and not torch.is_grad_enabled()
        rows = torch.arange(B, device=dev)
        # trunk columns of every head's first layer in one GEMM
        w_main = torch.cat([h.mlp_1.weight[:, :TRUNK] for h in heads], dim=0)
        pre_all = F.linear(main, w_main).float() if main.dtype != torch.float32 else F.linear(main, w_main)
        u = None
        if sampling and not deterministic:
            u = torch.rand((18, B), device=dev, generator=generator)
        out_cols: List[torch.Tensor] = [None] * L.ACTION_WORDS
        state = {"u": 0}

        def logits_of(h: int, extra: Optional[torch.Tensor]) -> torch.Tensor:
            hd = heads[h]
            x = pre_all[:, 128 * h:128 * (h + 1)] + hd.mlp_1.bias
            if extra is not None:
                x = x + F.linear(extra.to(hd.mlp_1.weight.dtype), hd.mlp_1.weight[:, TRUNK:])
            x = hd.mlp_2(F.relu(hd.norm(x)))
            return hd.distribution.linear(x).float()

        def categorical(logits, mask, given: Optional[torch.Tensor]):
            """-> (action [B] int64, log-prob [B], entropy [B]); ``mask`` float [B, D] of 0 / 1"""
            k = state["u"]
            state["u"] += 1
            if fused:
                from .policy_io import masked_categorical
                a, lp, ent = masked_categorical(logits.contiguous(), mask.float().contiguous(), deterministic=deterministic,
                                                uniforms=None if deterministic else u[k])
                return a.view(-1), lp.view(-1), ent
            logp_all = F.log_softmax(logits + torch.log(mask.float()), dim=-1)
            if given is not None:
                a = given
            elif deterministic:
                a = logp_all.argmax(dim=-1)
            else:                                                    # inverse CDF of the same uniform the fused kernel would take
                cdf = logp_all.exp().cumsum(-1)
                a = (cdf < (u[k] * cdf[:, -1]).unsqueeze(1)).sum(-1).clamp_(max=logits.shape[1] - 1)
                a = torch.where(mask.gather(1, a.view(-1, 1)).view(-1) > 0, a, logp_all.argmax(dim=-1))
            return a, logp_all.gather(1, a.view(-1, 1)).view(-1), _entropy(logp_all)

        def given(col: int) -> Optional[torch.Tensor]:
            return None if sampling else actions[:, col]

        joint = torch.zeros(B, device=dev)
        entropy = torch.zeros((), device=dev)

        def add(lp, ent, lpm):
            nonlocal joint, entropy
            if lpm is None:
                joint = joint + lp
                entropy = entropy + ent.mean()
            else:
                joint = joint + torch.where(lpm, lp, torch.zeros_like(lp))
                entropy = entropy + (ent * lpm).mean()

        # ---- head 0: action type
        a0, lp, ent = categorical(logits_of(0, None), masks[0], given(L.A_TYPE))
        add(lp, ent, None)
        typ = a0
        out_cols[L.A_TYPE] = a0
        is_ = lambda t: typ == t                                      # noqa: E731
        f = lambda *cols: torch.stack(cols, dim=-1).float()           # noqa: E731
        # ---- head 1: corner (type-conditional mask rows: settlement / city / dummy)
        m = masks[1][self._corner_row[typ], rows]
        a, lp, ent = categorical(logits_of(1, f(is_(PLACE_SETTLEMENT), is_(UPGRADE_CITY))), m, given(L.A_CORNER))
        add(lp, ent, is_(PLACE_SETTLEMENT) | is_(UPGRADE_CITY))
        out_cols[L.A_CORNER] = a
        # ---- heads 2, 3: edge, tile
        for h, col, t in ((2, L.A_EDGE, PLACE_ROAD), (3, L.A_TILE, MOVE_ROBBER)):
            a, lp, ent = categorical(logits_of(h, None), masks[h], given(col))
            add(lp, ent, is_(t))
            out_cols[col] = a
        # ---- head 4: development card
        card, lp, ent = categorical(logits_of(4, None), masks[4], given(L.A_CARD))
        add(lp, ent, is_(PLAY_DEV))
        out_cols[L.A_CARD] = card
        playing = is_(PLAY_DEV)
        # ---- head 5: accept / reject, with the proposed trade as custom input
        h5 = heads[5]
        cust = F.relu(h5.custom_norm(h5.custom_mlp(obs["proposed_trade"])))
        a, lp, ent = categorical(logits_of(5, cust), masks[5], given(L.A_ACCEPT))
        add(lp, ent, is_(RESPOND))
        out_cols[L.A_ACCEPT] = a
        # ---- head 6: player (rows: propose trade / steal / dummy)
        m = masks[6][self._player_row[typ], rows]
        a, lp, ent = categorical(logits_of(6, f(is_(PROPOSE), is_(STEAL))), m, given(L.A_PLAYER))
        add(lp, ent, is_(PROPOSE) | is_(STEAL))
        out_cols[L.A_PLAYER] = a

        # ---- heads 7, 8: recurrent resource pickers (action_heads_module.py:266-328)
        cur_res = obs["current_resources"].float()
        no_res = cur_res.sum(dim=-1) == 0
        proposing = is_(PROPOSE)

        def recurrent(h: int, base_extra: Optional[torch.Tensor], col0: int, by_hand: bool):
            hd = heads[h]
            x0 = pre_all[:, 128 * h:128 * (h + 1)] + hd.mlp_1.bias
            n_base = 0 if base_extra is None else base_extra.shape[1]
            if base_extra is not None:
                x0 = x0 + F.linear(base_extra.to(hd.mlp_1.weight.dtype), hd.mlp_1.weight[:, TRUNK:TRUNK + n_base])
            w_out = hd.mlp_1.weight[:, TRUNK + n_base:]
            output = torch.zeros((B, 6), device=dev)
            res = cur_res
            mask = (res > 0).float() if by_hand else torch.ones_like(res)
            mask = torch.cat((no_res.float().view(-1, 1), mask[:, 1:]), dim=1)       # "stop" first only for an empty hand
            lp_sum = torch.zeros(B, device=dev)
            ent_sum = torch.zeros(B, device=dev)
            prev = None
            for i in range(4):
                x = x0 + F.linear(output.to(w_out.dtype), w_out)
                logits = hd.distribution.linear(hd.mlp_2(F.relu(hd.norm(x)))).float()
                a, lp, ent = categorical(logits, mask, given(col0 + i))
                one_hot = F.one_hot(a, 6).float()
                if prev is not None:
                    live = prev > 0
                    lp, ent = torch.where(live, lp, torch.zeros_like(lp)), ent * live
                lp_sum, ent_sum = lp_sum + lp, ent_sum + ent
                output = output + one_hot
                res = (res - one_hot).clamp_(min=0)
                mask = (res > 0).float() if by_hand else torch.ones_like(res)
                mask = torch.cat((torch.ones((B, 1), device=dev), mask[:, 1:]), dim=1)
                output = torch.cat((torch.zeros((B, 1), device=dev), output[:, 1:]), dim=1)
                out_cols[col0 + i] = a
                prev = a
            lp_sum = torch.where(proposing, lp_sum, torch.zeros_like(lp_sum))
            ent_sum = ent_sum * proposing
            return output, lp_sum, ent_sum

        out7, lp7, ent7 = recurrent(7, None, L.A_GIVE, True)
        joint, entropy = joint + lp7, entropy + ent7.mean()
        filtered7 = (lp7 == 0).float().view(-1, 1)                    # action_heads_module.py:174
        out8, lp8, ent8 = recurrent(8, out7 * (1 - filtered7), L.A_RECV, False)
        joint, entropy = joint + lp8, entropy + ent8.mean()

        # ---- head 9: exchange / monopoly / year-of-plenty resource (two type-conditional mask rows multiplied)
        tflags = f(is_(PLAY_DEV), is_(EXCHANGE))
        cflags = f(card == CARD_YOP, card == CARD_MONOPOLY) * playing.float().view(-1, 1)        # filtered unless a card is played
        m_type = masks[9][self._exch_row_type[typ], rows]
        m_card = masks[9][self._exch_row_card[card], rows]
        m_card = torch.where(playing.view(-1, 1), m_card, torch.ones_like(m_card))
        a9, lp, ent = categorical(logits_of(9, torch.cat((tflags, cflags), dim=-1)), m_type * m_card, given(L.A_RES_A))
        lpm9 = (is_(PLAY_DEV) | is_(EXCHANGE)) & ((card == CARD_YOP) | (card == CARD_MONOPOLY) | ~playing)
        add(lp, ent, lpm9)
        out_cols[L.A_RES_A] = a9
        # ---- head 10: second resource
        extra = torch.cat((tflags, cflags, F.one_hot(a9, 5).float() * lpm9.float().view(-1, 1)), dim=-1)
        a, lp, ent = categorical(logits_of(10, extra), masks[10], given(L.A_RES_B))
        add(lp, ent, (is_(PLAY_DEV) | is_(EXCHANGE)) & ((card == CARD_YOP) | ~playing))
        out_cols[L.A_RES_B] = a
        # ---- head 11: discard
        a, lp, ent = categorical(logits_of(11, None), masks[11], given(L.A_DISCARD))
        add(lp, ent, is_(DISCARD))
        out_cols[L.A_DISCARD] = a

        zero = torch.zeros(B, dtype=torch.int64, device=dev)
        action_rows = torch.stack([zero if c is None else c for c in out_cols], dim=1)
        return action_rows, joint.view(-1, 1), entropy

    # ------------------------------------------------------------------ public surface (policy.py:71-111)
    @torch.no_grad()
    def act(self, obs, masks, deterministic: bool = False, generator: Optional[torch.Generator] = None):
        """-> (normalised value [B, 1], action rows int32 [B, 20], joint log-prob fp32 [B, 1])"""
        value, main = self.base(obs)
        rows, logp, _ = self._run_heads(main, obs, masks, None, deterministic, generator)
        return value.float(), rows.to(torch.int32), logp

    def evaluate_actions(self, obs, masks, action_rows: torch.Tensor):
        """-> (normalised value [B, 1], joint log-prob [B, 1], entropy scalar) of stored int32 / int64 [B, 20] action rows"""
        value, main = self.base(obs)
        _, logp, entropy = self._run_heads(main, obs, masks, action_rows.long(), False, None)
        return value.float(), logp, entropy
