"""Timeline of CTA 0 of the f16 node GEMM (5 planes + q MLP, 162 row tiles = config 2) from %globaltimer stamps
(cbg_debug_node_gemm_trace), plus the launch time by CUDA events.  Usage: python scripts/trace_node_gemm.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from cbgbench_b200 import _lib  # noqa: E402
from helpers import make_model  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    model, _ = make_model(10, device=dev)
    L = _lib.lib()
    lay = _lib.blob_layout()
    blob = model.denoiser.packed_blob(dev)
    N = 20736
    h = torch.randn(N, 128, device=dev)
    planes = torch.zeros(5, N, 128, device=dev)
    base_ptr = blob.data_ptr() + 4 * (lay['global_floats'] + 3 * lay['layer_floats'])
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for sub in (0, 1):
        for it in range(4):
            _lib.check(L.cbg_node_proj_f32(base_ptr, sub, 2, h.data_ptr(), None, N, N, planes.data_ptr(), None))
        torch.cuda.synchronize()
        ts = []
        for it in range(10):
            flush.zero_()
            h.add_(0.0)                                    # h back into L2, as after the aggregation kernel
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.cbg_node_proj_f32(base_ptr, sub, 2, h.data_ptr(), None, N, N, planes.data_ptr(), None))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f'sublayer {sub}: 5 planes + q, {N} rows: launch {np.median(ts):.1f} us (min {min(ts):.1f})')
    _lib.check(L.cbg_debug_node_gemm_trace(buf.data_ptr()))
    _lib.check(L.cbg_node_proj_f32(base_ptr, 0, 2, h.data_ptr(), None, N, N, planes.data_ptr(), None))
    torch.cuda.synchronize()
    _lib.check(L.cbg_debug_node_gemm_trace(None))
    t = buf.cpu().numpy()
    t0 = t[0]
    print('CTA 0 (us from start): A staged', (t[1] - t0) / 1e3)
    for g in range(6):
        print(f'  gemm {g}: accumulator complete {(t[2 + g] - t0) / 1e3:7.2f}   epilogue done {(t[10 + g] - t0) / 1e3:7.2f}')
    print('  end', (t[20] - t0) / 1e3)


if __name__ == '__main__':
    main()
