"""Sinusoidal positional table with period `period` (PeriodicPositionalEncoding buffer `pe`,
/root/reference/models/emage_audio/processing_emage_audio.py:328-340): sin on even columns, cos on odd
columns, frequencies 10000^(-2i/d).  Computed in float32 torch ops so the buffer is bit-identical to
the one the reference registers (it is also stored in every checkpoint)."""
import math

import torch


def periodic_table(d_model: int, period: int) -> torch.Tensor:
    pos = torch.arange(period, dtype=torch.float32).unsqueeze(1)
    freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
    table = torch.zeros(period, d_model)
    table[:, 0::2] = torch.sin(pos * freq)
    table[:, 1::2] = torch.cos(pos * freq)
    return table
