"""Issue rates of the DP kernels' instructions on this GPU, alone and interleaved in pairs (csrc/exp_issue.cu).

    python scripts/issue_rates.py [out.json]

Prints one line per case: thread-level instructions per second and the same per SM and clock (lanes / clk / SM).
"""
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from nextgenmap_b200.host import CudaSW  # noqa: E402

NAMES = ["VIADDMNMX.S16x2", "IMAD", "VIMNMX3.S16x2", "PRMT", "VIADD.16x2", "LOP3", "SHF.R.W", "VIMNMX.S16x2", "VIADDMNMX.U16x2", "VIADDMNMX.S16x2.RELU", "IADD (add.u32)"]


def main():
    sw = CudaSW(152, 27)
    lib = sw.lib
    lib.ngm_b200_exp_issue_rates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    cap = 64
    a, b, r = (C.c_int * cap)(), (C.c_int * cap)(), (C.c_double * cap)()
    n = lib.ngm_b200_exp_issue_rates(sw.ctx, cap, a, b, r)
    assert n > 0, sw._err()
    props = torch.cuda.get_device_properties(0)
    sms = props.multi_processor_count
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
    rows = []
    mixes = ["slot mix 0: as built (VIADD.16x2, VIADDMNMX, VIADDMNMX.RELU, LOP3, PRMT, 2 IMAD)", "slot mix 1: u de-fused (2 VIADD.16x2, VIMNMX, VIADDMNMX.RELU, ...)",
             "slot mix 2: all de-fused (3 VIADD.16x2, 3 VIMNMX, ...)", "slot mix 3: mix 0 + VIMNMX3 per two slots", "slot mix 4: mix 0 + VIMNMX per slot"]
    for i in range(n):
        name = mixes[a[i] - 100] if a[i] >= 100 else NAMES[a[i]] + (" + " + NAMES[b[i]] if b[i] >= 0 else "")
        lanes = r[i] / (sms * mhz * 1e6)
        rows.append({"case": name, "thread_instr_per_s": r[i], "lanes_per_clk_per_sm_at_max_clock": lanes})
        print(f"{name:44s} {r[i] / 1e12:7.2f} T/s   {lanes:6.1f} lanes/clk/SM" + (f"   = {64.0 / lanes:5.2f} ALU-slot equivalents per slot" if a[i] >= 100 else ""))
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(json.dumps({"sms": sms, "sm_max_mhz": mhz, "cases": rows}, indent=1))
    sw.close()


if __name__ == "__main__":
    main()
