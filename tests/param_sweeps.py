"""Parameter sweeps shared by tests/test_gpu_z_params.py (CUDA path vs oracle) and tests/test_params_cpu.py (oracle vs the
compiled reference): SURVEY 8d's non-default settings plus the setters' guards."""
import math

from hackrfdiags_b200 import capi

RX_MODES = [capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM, capi.MODE_LSB, capi.MODE_USB]
NAMES = {1: "am", 2: "fm", 3: "wbfm", 4: "lsb", 5: "usb"}
GAIN_PARAM = {capi.MODE_AM: capi.PARAM_AM_GAIN, capi.MODE_FM: capi.PARAM_FM_GAIN, capi.MODE_WBFM: capi.PARAM_WBFM_GAIN,
              capi.MODE_LSB: capi.PARAM_SSB_GAIN, capi.MODE_USB: capi.PARAM_SSB_GAIN}
# the demodulators' default gains (AmDemodulator.cc:102, FmDemodulator.cc:173, WbFmDemodulator.cc:151,
# SsbDemodulator.cc:146); the GPU test reads them back through hrd_get_param, the CPU test uses this copy
DEFAULT_GAIN = {capi.MODE_AM: 300.0, capi.MODE_FM: 64000 / (2 * math.pi), capi.MODE_WBFM: 256000 / (2 * math.pi),
                capi.MODE_LSB: 300.0, capi.MODE_USB: 300.0}


def rx_gain_sweep(default: float):
    """One entry per stream; None = leave the default.  x0.1 / x10, tiny, huge (int16 wrap), zero, negative."""
    return [None, default * 0.1, default * 10.0, 30.0, 30000.0, 0.0, -default, 1.0]


def tx_setter_sweeps(mode: int):
    """One LIST of setter calls per stream, accepted and rejected ones mixed."""
    if mode == capi.MODE_AM:  # accepted iff 0 <= m <= 1
        return [[], [0.3], [1.0], [0.0], [1.5], [0.3, -0.1], [0.75, 1.0]]
    if mode == capi.MODE_FM:  # the guard tests the OLD value against 0..3500
        return [[], [1000.0], [3500.0], [9000.0], [9000.0, 100.0], [0.0], [1000.0, 2500.0], [-500.0, 1000.0]]
    if mode == capi.MODE_WBFM:  # the same against 0..112000
        return [[], [1000.0], [112000.0], [50000.0], [200000.0], [200000.0, 1000.0], [0.0], [-500.0, 1000.0]]
    raise ValueError(mode)
