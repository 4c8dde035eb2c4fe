#!/usr/bin/env python3
"""Dev tool (GPU box): one DCPreESR pass over B x T samples (for ncu captures).  usage: esr_once.py B T"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ntm_b200

B, T = int(sys.argv[1]), int(sys.argv[2])
t = 0.2 * torch.randn(B, 1, T, device="cuda:0")
o = t + 0.02 * torch.randn(B, 1, T, device="cuda:0")
print(float(ntm_b200.DCPreESR()(o, t)), float(ntm_b200.DCPreESR()(o, t)))
