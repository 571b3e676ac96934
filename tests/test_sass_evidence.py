"""CPU: the shipped library is sm_100a-only code and contains the Blackwell instructions DESIGN.md claims —
tcgen05 (UTCHMMA / UTCBAR / LDTM) in the head contraction, bulk TMA + mbarrier (UBLKCP / SYNCS) in the
staged kernels, packed fp32x2 math (FFMA2) and 3-input min/max (FMNMX3) in lift+argmax
(mnemonics: /opt/skills/guides/B200_PROFILING.md)."""
import re
import shutil
import subprocess

import pytest


@pytest.fixture(scope="module")
def sass(built_lib):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        elfs = subprocess.run([exe, "-lelf", built_lib], capture_output=True, text=True, timeout=120).stdout
        text = subprocess.run([exe, "-sass", built_lib], capture_output=True, text=True, timeout=600).stdout
    except (FileNotFoundError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    return elfs, text


def test_every_cubin_is_sm_100a(sass):
    elfs, _ = sass
    names = re.findall(r"ELF file\s+\d+:\s+(\S+)", elfs)
    assert len(names) >= 9 and all(n.endswith(".sm_100a.cubin") for n in names), names


def test_blackwell_instructions_are_present(sass):
    _, text = sass

    keys = ("UTCHMMA", "UTCBAR", "LDTM", r"UBLKCP\.S\.G", r"SYNCS\.PHASECHK", "FFMA2", "FMNMX3", r"MUFU\.LG2",
            r"ST[G]?\.E\.64\.STRONG\.SYS", r"UTMALDG\.[23]D", r"MUFU\.RCP", r"LDGSTS")
    rx = re.compile(r"\b(?:" + "|".join(f"(?P<g{i}>{k})" for i, k in enumerate(keys)) + ")")
    found = {k: set() for k in keys}
    cur = None
    for line in text.splitlines():                      # one pass: which functions contain which mnemonic
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
        elif cur:
            m = rx.search(line)
            if m:
                found[keys[int(m.lastgroup[1:])]].add(cur)

    def per_function(mnemonic):
        return found[mnemonic]

    tc = per_function("UTCHMMA")
    heads = ("head_logits_tc_kernel", "head_logits_tma_kernel", "cell_classify_tma_kernel")
    assert tc and all(any(h in f for h in heads) for f in tc)                   # tensor cores only in the two classifier heads
    assert all(any(h in f for f in tc) for h in heads)
    tma = per_function(r"UTMALDG\.[23]D")                                       # tensor-map TMA feeds both tcgen05 contractions
    assert tma and all("head_logits_tma_kernel" in f or "cell_classify_tma_kernel" in f for f in tma) and tma <= tc
    assert any("cell_classify_tma_kernel" in f for f in tma)
    assert per_function("UTCBAR") and per_function("LDTM")                      # tcgen05.commit, tcgen05.ld
    bulk = per_function(r"UBLKCP\.S\.G")
    assert any("lift_separable_kernel" in f for f in bulk) and any("decode_tail_tma_kernel" in f for f in bulk)
    assert per_function(r"SYNCS\.PHASECHK") >= bulk                             # every bulk copy is waited for by mbarrier
    assert any("lift_argmax" in f for f in per_function("FFMA2")) and any("lift_argmax" in f for f in per_function("FMNMX3"))
    lg2 = per_function(r"MUFU\.LG2")
    assert any("laplace_qsample_kernel" in f for f in lg2) and any("laplace_qsample_map_kernel" in f for f in lg2)
    push = per_function(r"ST[G]?\.E\.64\.STRONG\.SYS")                        # the peer push rides in all three producers
    assert all(any(k in f for f in push) for k in ("confusion_hist", "lut_paint_hist", "lift_argmax_env", "xchg_push"))
    env = [f for f in per_function("FFMA2") if "lift_argmax_env" in f]
    assert env and any("lift_argmax_env" in f for f in per_function(r"MUFU\.RCP"))   # the envelope sweep's run lengths
    assert any("lift_argmax_env" in f for f in per_function("LDGSTS"))          # cp.async ground-truth tile
    row = [f for f in per_function("FFMA2") if "lift_argmax_row" in f]          # the row form: same packed sweep
    assert row and any("lift_argmax_row" in f for f in per_function(r"MUFU\.RCP"))
    assert any("lift_argmax_row" in f for f in per_function("FMNMX3"))
