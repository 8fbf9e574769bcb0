"""CPU-side checks: the C-ABI library loads and exports every symbol include/vdqn.h declares,
the ctypes structs match the header, the drop-in module keeps the reference's layout, and the
product path refuses to run without a GPU (no fallback)."""
import ctypes
import json
import os
import re
import subprocess
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "vdqn.h")


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    so = os.path.join(ROOT, "video_dqn_b200", "libvdqn.so")
    if not os.path.exists(so):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "video_dqn_b200", "build.py")])
    from video_dqn_b200 import _lib
    return _lib


def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vdqn_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 19
    h = lib.load()
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/vdqn.h but not exported by libvdqn.so"
        assert n in lib.EXPORTS, f"{n} has no ctypes signature"
    assert h.vdqn_abi_version() == 1


def test_struct_sizes_match_header(lib):
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirrors."""
    src = '#include <stdio.h>\n#include "vdqn.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",' \
          "sizeof(vdqn_conv_desc),sizeof(vdqn_wgrad_desc),sizeof(vdqn_wgrad_fin_desc)," \
          "sizeof(vdqn_wprep_desc),sizeof(vdqn_td_desc));return 0;}"
    exe = "/tmp/vdqn_sizes"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=src.encode(), check=True)
    sizes = list(map(int, subprocess.check_output([exe]).split()))
    mine = [ctypes.sizeof(c) for c in (lib.ConvDesc, lib.WgradDesc, lib.WgradFinDesc, lib.WprepDesc, lib.TdDesc)]
    assert sizes == mine


def test_no_gpu_means_loud_failure(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = lib.load()
    assert h.vdqn_init(0) != 0
    assert b"CUDA" in h.vdqn_last_error() or b"cuda" in h.vdqn_last_error()


def test_module_keeps_reference_layout():
    sys.path.insert(0, ROOT)
    from video_dqn_b200.qnet import HabitatDQNMultiAction, grad_param_names
    lay = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_module_layout.json")))
    m = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False)
    sd = m.state_dict()
    assert list(sd.keys()) == lay["state_dict_keys"] and len(sd) == 250
    assert [n for n, _ in m.named_parameters()] == lay["named_parameters"]
    assert len(list(m.parameters())) == 70
    for k, v in sd.items():
        assert list(v.shape) == lay["shapes"][k], k
    names = grad_param_names()
    assert len(names) == 68
    order = [n for n, _ in m.named_parameters()]
    idx = [order.index(n) for n in names]
    assert idx == sorted(idx) and idx[:60] == list(range(60)) and idx[60:] == list(range(62, 70))
    assert sum(dict(m.named_parameters())[n].numel() for n in names) == 12426383
    # aliasing: features.0-7 share storage with resnet.*
    assert sd["features.0.weight"].data_ptr() == sd["resnet.conv1.weight"].data_ptr()
    assert m.num_frames == 1 and HabitatDQNMultiAction(3).num_frames == 4
    # set_train keeps the trunk's BatchNorm in eval mode
    m.set_train()
    assert m.training and not m.resnet.bn1.training and m.top.training
    with pytest.raises(Exception, match="bad shape"):
        m(torch.zeros(2, 4, 3, 224, 224))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 224, 224))


def test_plan_matches_survey_flop_count():
    sys.path.insert(0, ROOT)
    from video_dqn_b200 import engine as E
    plan = E.make_plan(3)
    assert len(plan.convs) == 21
    # algorithmic MACs per frame of the *reference* network (7x7x3 stem), SURVEY.md 8: 1 821 888 256
    macs = 112 * 112 * 64 * 147
    for c in plan.convs[1:]:
        macs += c.out_hw * c.out_hw * c.cout * c.cin * c.k * c.k
    macs += 1600 * 512 + 512 * 256 + 256 * 15
    assert macs == 1821888256
    for c in plan.convs:
        assert 1 <= E.wgrad_splits(c, 256, sms=148) <= 4096


def test_inverse_module_container_keeps_reference_layout():
    """`InverseActionModule` (the parameter container behind the inverse-model runner / trainer) has the
    key layout that oracle.inverse.init_state was loaded into the reference's own module with
    (oracle/make_inverse_goldens.py, strict), trains exactly the reference's 12 head tensors and has no
    CPU forward."""
    sys.path.insert(0, ROOT)
    from oracle import inverse as oinv
    from video_dqn_b200.inverse import InverseActionModule
    m = InverseActionModule()
    ref = oinv.init_state(seed=7)
    assert set(m.state_dict()) == set(ref)
    assert all(m.state_dict()[k].shape == v.shape for k, v in ref.items())
    assert tuple(n for n, p in m.named_parameters() if p.requires_grad) == oinv.TRAINABLE
    assert not m.resnet18.training
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 224, 224), torch.zeros(1, 3, 224, 224))


def test_new_wrappers_refuse_cpu_tensors_before_launch(lib):
    """No CPU fallback anywhere: the tensor-level wrappers of the inverse-model / train-mode-BatchNorm
    kernels raise on host tensors before the library is called."""
    from video_dqn_b200 import ops
    x = torch.zeros(2, 4, 4, 64, dtype=torch.bfloat16)
    c = torch.zeros(64)
    with pytest.raises(ValueError, match="CUDA"):
        ops.cross_entropy(torch.zeros(4, 3), torch.zeros(4, dtype=torch.int64))
    with pytest.raises(ValueError, match="CUDA"):
        ops.dropout_mask(torch.zeros(8, dtype=torch.uint8), 0.5, 0, 0)
    with pytest.raises(ValueError, match="CUDA"):
        ops.dropout_apply(torch.zeros(8), torch.zeros(8, dtype=torch.uint8), 2.0)
    with pytest.raises(ValueError, match="CUDA"):
        ops.avgpool_bwd(torch.zeros(2, 64), x)
    st = types.SimpleNamespace(C=64)
    with pytest.raises(ValueError, match="CUDA"):
        ops.bn_train_fwd(x, st, c, c, c, c, None, x.clone())
    with pytest.raises(ValueError, match="CUDA"):
        ops.bn_train_bwd(x, x, st, c, c, c, x.clone())
    from video_dqn_b200.inverse import InverseActionModule, InverseModelTrainer
    with pytest.raises(RuntimeError, match="no CPU path"):
        InverseModelTrainer(InverseActionModule().state_dict(), 4, device="cpu")
