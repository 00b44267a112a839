set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02x; mkdir -p $O
# Welch at 8 points per thread, 512 threads, 64 registers (2 CTAs/SM = 32 warps/SM) against 16 points per thread (16 warps/SM)
timeout 300 python tests/tools/sweep.py --no-cpu --ratios 1 --sizes 2048,4096,8192 > $O/sweep_r1_ppt16.log 2>&1
timeout 300 python tests/tools/sweep.py --no-cpu --ratios 1 --sizes 2048,4096,8192 --lib pypanadapter_b200/libzoomfft_alt_ppt8.so > $O/sweep_r1_ppt8.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --lib pypanadapter_b200/libzoomfft_alt_ppt8.so > $O/bench_cfg2_ppt8.json 2> $O/bench_cfg2_ppt8.err
# virtual receivers with few channels per GPU (what an 8-GPU cfg4 rank sees): host time per call
timeout 300 python - > $O/cfg4_8ch.log 2>&1 <<'PY'
import json, numpy as np, torch
from pypanadapter_b200 import synth
from pypanadapter_b200.engine import ZoomPSD
w = synth.WORKLOADS["cfg4"]
F = 8
host = synth.make_frames(w, F, distinct=F)
d_in = torch.from_numpy(host.view(np.uint8).reshape(F, -1)).cuda()
eng = ZoomPSD(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
eng.configure(w.fs, w.fft_size, w.fft_ratio, w.frame_len, w.window, dtype=w.dtype, f_demod=w.f_demod, crop=w.crop)
for nch in (8, 64):
    centres = synth.cfg4_centres()[:nch]
    rows = torch.empty((nch * F, eng.row_width), dtype=torch.float32, device="cuda")
    for _ in range(5): eng.process_channels_device(d_in.data_ptr(), F, centres, rows.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50): eng.process_channels_device(d_in.data_ptr(), F, centres, rows.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(json.dumps({"channels": nch, "frames": F, "ms_per_step": ms, "g_channel_samples_per_s": nch * F * w.frame_len / ms / 1e6}))
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg4 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
ls -la $O
