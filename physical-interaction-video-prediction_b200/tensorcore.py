"""bf16 tensor-core plan for the seven ConvLSTM layers: bf16 operand shadows, prepared weights, tcgen05 launches.

Created by ``Engine`` when ``compute="bf16"``.  fp32 stays the master format of every activation that the
bandwidth-bound kernels consume (cell state, LayerNorm input, gradients); this class only adds the bf16 GEMM operands:
``xh_bf16[l][t]`` (concatenated layer input, channels padded to a multiple of 64), ``dg_bf16[l]`` (gate pre-activation
gradients) and the K-major weight copies ``Wf`` / ``Wd`` refreshed after every Adam step.
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
        self.Kpad = [(cin + c + 63) // 64 * 64 for cin, c in zip(LSTM_IN, LSTM_SIZES)]
        self.xh_bf16, self.dg_bf16, self.Wf, self.Wd = [], [], [], []
        for li, (cin, c, lv) in enumerate(zip(LSTM_IN, LSTM_SIZES, LSTM_LEVEL)):
            M = ws["Mr"][lv]
            if M % 128:
                raise PivpError("bf16 tensor-core path: B*H*W = %d of ConvLSTM layer %d is not a multiple of 128 "
                                "(use an even batch / larger images, or compute='f32')" % (M, li + 1))
            self.xh_bf16.append([torch.zeros(M, self.Kpad[li], dtype=torch.bfloat16, device=dev) for _ in range(T)])
            self.dg_bf16.append(torch.empty(M, 4 * c, dtype=torch.bfloat16, device=dev))
            self.Wf.append(torch.empty(4 * c, 25, self.Kpad[li], dtype=torch.bfloat16, device=dev))
            self.Wd.append(torch.empty(cin + c, 25, 4 * c, dtype=torch.bfloat16, device=dev))
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
        e, ws = self.eng, self.ws
        cin, C, lv = LSTM_IN[li], LSTM_SIZES[li], LSTM_LEVEL[li]
        h, w = e.H // lv, e.W // lv
        cx = cin + C
        e.L.call("pivp_tc_conv5x5", _ptr(self.dg_bf16[li]), 4 * C, ws["B"], h, w, 4 * C,
                 _ptr(self.Wd[li]), cx, cx, 0, 0,
                 _ptr(ws["dxh"][li]), cx, 0,
                 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                 C, 0.0, 0, e._s())
