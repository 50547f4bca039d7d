"""bf16 tensor-core plan for the seven ConvLSTM layers: bf16 operand shadows, prepared weights, tcgen05 launches.

Created by ``Engine`` when ``compute="bf16"``.  fp32 stays the master format of every activation that the
bandwidth-bound kernels consume (cell state, LayerNorm input, gradients); this class only adds the bf16 GEMM operands:
  xh_bf16[l][t]  (M, Kpad)   concatenated layer input [x | h_{t-1}], channels padded to a multiple of 64 (fwd GEMM A operand)
  dg_bf16[l][t]  (M, 4C)     gate pre-activation gradients (input-gradient GEMM A operand)
  (the weight-gradient GEMM reads xh_bf16 / dg_bf16 of ALL time steps MN-major, straight from these NHWC tensors)
  Wf[l], Wd[l]               K-major bf16 weights (forward / tap-flipped input-gradient), refreshed after every Adam step
"""
import torch

from .engine import LSTM_IN, LSTM_SIZES, LSTM_LEVEL, View, _ptr
from ._lib import PivpError


class TensorCorePlan(object):
    def __init__(self, eng, ws):
        self.eng = eng
        self.ws = ws
        dev = eng.dev
        B, T = ws["B"], ws["T"]
        S = T - 1
        self.S = S
        self.Kpad = [(cin + c + 63) // 64 * 64 for cin, c in zip(LSTM_IN, LSTM_SIZES)]
        self.xh_all, self.dg_all, self.xh_bf16, self.dg_bf16 = [], [], [], []
        self.Wf, self.Wd = [], []
        wsb = 16
        for li, (cin, c, lv) in enumerate(zip(LSTM_IN, LSTM_SIZES, LSTM_LEVEL)):
            M = ws["Mr"][lv]
            h, w = eng.H // lv, eng.W // lv
            if M % 128 or (h * w) % 64:
                raise PivpError("bf16 tensor-core path: ConvLSTM layer %d has B*H*W = %d (needs a multiple of 128) and H*W = %d "
                                "(needs a multiple of 64); use an even batch / 64x64 or larger images, or compute='f32'"
                                % (li + 1, M, h * w))
            xa = torch.zeros(T, M, self.Kpad[li], dtype=torch.bfloat16, device=dev)      # zeros: h_{-1}, pad channels
            da = torch.empty(S, M, 4 * c, dtype=torch.bfloat16, device=dev)
            self.xh_all.append(xa)
            self.dg_all.append(da)
            self.xh_bf16.append([xa[t] for t in range(T)])
            self.dg_bf16.append([da[t] for t in range(S)])
            self.Wf.append(torch.empty(4 * c, 25, self.Kpad[li], dtype=torch.bfloat16, device=dev))
            self.Wd.append(torch.empty(cin + c, 25, 4 * c, dtype=torch.bfloat16, device=dev))
            wsb = max(wsb, eng.L.query("pivp_tc_wgrad_workspace_bytes", S * B, h, w, cin + c, 4 * c))
        self.wgrad_ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        self.accurate = 0
        self.refresh_weights()

    def refresh_weights(self):
        e = self.eng
        for li, (cin, c) in enumerate(zip(LSTM_IN, LSTM_SIZES)):
            e.L.call("pivp_tc_prep_weights", _ptr(e.p["lstm%d/conv/W" % (li + 1)]), 4 * c, cin + c, self.Kpad[li],
                     _ptr(self.Wf[li]), _ptr(self.Wd[li]), e._s())

    def xview(self, li, t):
        """bf16 x-slot of layer li at time t (what the producer of the layer input also writes)."""
        return View(self.xh_bf16[li][t], self.Kpad[li], 0, LSTM_IN[li])

    def lstm_fwd(self, li, t):
        """conv + bias + gates + cell + h of BasicConvLSTMCell (train_model.py:262-272) in ONE tcgen05 kernel."""
        e, ws = self.eng, self.ws
        cin, C, lv = LSTM_IN[li], LSTM_SIZES[li], LSTM_LEVEL[li]
        h, w = e.H // lv, e.W // lv
        e.L.call("pivp_tc_conv5x5", _ptr(self.xh_bf16[li][t]), self.Kpad[li], ws["B"], h, w, self.Kpad[li],
                 _ptr(self.Wf[li]), 4 * C, 128, 1, _ptr(e.p["lstm%d/conv/b" % (li + 1)]),
                 0, 0, 0,
                 _ptr(ws["G"][li][t]), _ptr(ws["c"][li][t - 1]) if t > 0 else 0, _ptr(ws["c"][li][t]),
                 _ptr(ws["xh"][li][t + 1]), cin + C, cin, _ptr(self.xh_bf16[li][t + 1]), self.Kpad[li], cin,
                 0, 0, 0,
                 C, 1.0, self.accurate, e._s())

    def lstm_dgrad(self, li, t):
        """dxh = conv(dG_t, tap-flipped W): gradient w.r.t. the concatenated input [x | h_{t-1}] (D.5)."""
        e, ws = self.eng, self.ws
        cin, C, lv = LSTM_IN[li], LSTM_SIZES[li], LSTM_LEVEL[li]
        h, w = e.H // lv, e.W // lv
        cx = cin + C
        e.L.call("pivp_tc_conv5x5", _ptr(self.dg_bf16[li][t]), 4 * C, ws["B"], h, w, 4 * C,
                 _ptr(self.Wd[li]), cx, cx, 0, 0,
                 _ptr(ws["dxh"][li]), cx, 0,
                 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                 C, 0.0, 0, e._s())

    def wgrad_all(self):
        """After BPTT: weight and bias gradients of all seven ConvLSTM convolutions, each as ONE GEMM over all time steps."""
        e, ws = self.eng, self.ws
        S, B = self.S, ws["B"]
        for li, (cin, C, lv) in enumerate(zip(LSTM_IN, LSTM_SIZES, LSTM_LEVEL)):
            M = ws["Mr"][lv]
            h, w = e.H // lv, e.W // lv
            cx = cin + C
            name = "lstm%d/conv" % (li + 1)
            e.L.call("pivp_tc_colsum_bf16", _ptr(self.dg_all[li]), 4 * C, S * M, 4 * C, _ptr(e.g[name + "/b"]), e._s())
            e.L.call("pivp_tc_wgrad5x5", _ptr(self.dg_all[li]), _ptr(self.xh_all[li]), self.Kpad[li], S * B, h, w, cx, 4 * C,
                     _ptr(e.g[name + "/W"]), _ptr(self.wgrad_ws), self.wgrad_ws.numel(), e._s())
