"""Oracle restatement of transformers' CLIPTextModel as configured for Stable Diffusion 1.5 (openai/clip-vit-large-patch14
text tower) — TEST INFRASTRUCTURE ONLY. SURVEY.md 8(a) row a13 / 8(f) row 3.

The reference calls it at diffusert/lcm/lcm_controlnet.py:175-179:
    prompt_embeds = self.text_encoder(text_input_ids.to(device), attention_mask=None)[0]      # last_hidden_state
Config [transformers-knowledge]: vocab 49408, 77 positions, hidden 768, 12 layers, 12 heads (d = 64), MLP 3072,
hidden_act "quick_gelu" (x * sigmoid(1.702 x)), layer_norm_eps 1e-5, causal attention mask, final_layer_norm.
123 060 480 parameters. Module / parameter names reproduce the transformers state-dict keys so real checkpoints load.

Pinned: tests/test_oracle.py loads the same seeded state dict into transformers.CLIPTextModel (installed in this image,
v5.5) and requires equality to 1e-5 — this piece of the oracle is NOT "parity unpinned".
"""
import torch
import torch.nn as nn

VOCAB, POSITIONS, HIDDEN, LAYERS, HEADS, MLP = 49408, 77, 768, 12, 12, 3072


class _Attention(nn.Module):
    def __init__(self):
        super().__init__()
        self.q_proj = nn.Linear(HIDDEN, HIDDEN)
        self.k_proj = nn.Linear(HIDDEN, HIDDEN)
        self.v_proj = nn.Linear(HIDDEN, HIDDEN)
        self.out_proj = nn.Linear(HIDDEN, HIDDEN)

    def forward(self, x, mask):
        b, t, c = x.shape
        d = c // HEADS
        q = (self.q_proj(x) * d ** -0.5).view(b, t, HEADS, d).transpose(1, 2)
        k = self.k_proj(x).view(b, t, HEADS, d).transpose(1, 2)
        v = self.v_proj(x).view(b, t, HEADS, d).transpose(1, 2)
        s = q @ k.transpose(-1, -2) + mask
        o = torch.softmax(s, dim=-1) @ v
        return self.out_proj(o.transpose(1, 2).reshape(b, t, c))


class _MLP(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(HIDDEN, MLP)
        self.fc2 = nn.Linear(MLP, HIDDEN)

    def forward(self, x):
        h = self.fc1(x)
        return self.fc2(h * torch.sigmoid(1.702 * h))


class _Layer(nn.Module):
    def __init__(self):
        super().__init__()
        self.self_attn = _Attention()
        self.layer_norm1 = nn.LayerNorm(HIDDEN, eps=1e-5)
        self.mlp = _MLP()
        self.layer_norm2 = nn.LayerNorm(HIDDEN, eps=1e-5)

    def forward(self, x, mask):
        x = x + self.self_attn(self.layer_norm1(x), mask)
        return x + self.mlp(self.layer_norm2(x))


class _Embeddings(nn.Module):
    def __init__(self):
        super().__init__()
        self.token_embedding = nn.Embedding(VOCAB, HIDDEN)
        self.position_embedding = nn.Embedding(POSITIONS, HIDDEN)


class _Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([_Layer() for _ in range(LAYERS)])


class _TextModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.embeddings = _Embeddings()
        self.encoder = _Encoder()
        self.final_layer_norm = nn.LayerNorm(HIDDEN, eps=1e-5)


class ClipTextOracle(nn.Module):
    def __init__(self):
        super().__init__()
        self.text_model = _TextModel()

    @torch.no_grad()
    def forward(self, input_ids):
        """input_ids: int64 (B, T<=77) -> last_hidden_state fp32 (B, T, 768)."""
        tm = self.text_model
        b, t = input_ids.shape
        x = tm.embeddings.token_embedding(input_ids) + tm.embeddings.position_embedding(torch.arange(t))[None]
        mask = torch.full((t, t), float("-inf")).triu(1)[None, None]
        for layer in tm.encoder.layers:
            x = layer(x, mask)
        return tm.final_layer_norm(x)
