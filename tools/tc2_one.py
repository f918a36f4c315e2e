"""Runs the fused SRCNN kernel a few times on a 3840x2160 Y plane (profiling target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srcnn_cpp_b200 as S
st = torch.cuda.Stream()
torch.cuda.set_stream(st)   # torch and the context on one stream
eng = S.Engine(0, stream=st.cuda_stream)
y = torch.randint(0, 256, (2160, 3840), dtype=torch.uint8, device="cuda")
out = torch.zeros_like(y)
for _ in range(6):
    eng.stage_cnn(y, out)
eng.sync()
