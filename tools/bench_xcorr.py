#!/usr/bin/env python
"""Matched-filter (gf3_xcorr) throughput for the two chirp lengths that matter: 21 600 samples (N = 4096
modes, 11 partitions of 2048) and 5 280 samples (the C3 geometry, 3 partitions).
usage (from the repo root): python tools/bench_xcorr.py"""
import os, sys; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gf3-audio-modem_b200"))
import torch, gf3b200
for cfg,B,T in ((dict(N=4096,cp=224,lo=100,hi=1500),64,993600+21600),(dict(N=1024,cp=32,lo=1,hi=512),1024,242984)):
    phy=gf3b200.Phy(**cfg)
    r=torch.randn((B,T),device="cuda")
    for _ in range(2): P,pm=phy.xcorr(r)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): P,pm=phy.xcorr(r)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(cfg["N"],"chirp_len",phy.chirp_len,"parts",(phy.chirp_len+2047)//2048,"streams",B,"T",T,"xcorr %.3f ms"%ms,"-> %.1f Gsample/s"%(B*T/ms/1e6))
