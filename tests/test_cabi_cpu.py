"""CPU: the C-ABI library loads, exports every symbol include/tfmq_b200.h declares, the ctypes
structures match the C layouts, and contexts fail loudly without an sm_100 device."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tfmq_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfmq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tfmq_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding and header disagree"
    assert lib.tfmq_abi_version() == 6


def test_struct_layouts_match_c():
    from tfmq_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "tfmq_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(tfmq_act_desc), sizeof(tfmq_conv_w4a8_desc), sizeof(tfmq_conv_fp_desc),
         sizeof(tfmq_linear_desc), sizeof(tfmq_attn_desc), sizeof(tfmq_conv_h16_desc), sizeof(tfmq_attn_h16_desc));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [ctypes.sizeof(t) for t in (_lib.ActDesc, _lib.ConvW4A8Desc, _lib.ConvFpDesc, _lib.LinearDesc, _lib.AttnDesc,
                                     _lib.ConvH16Desc, _lib.AttnH16Desc)]
    assert sizes == want


def test_context_fails_loudly_without_gpu():
    import torch
    from tfmq_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="tfmq_create failed"):
        _lib.Context(0)


def test_ops_refuse_cpu_tensors():
    import torch
    from tfmq_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.gn_stats(torch.zeros(1, 4, 4, 32), 32)
