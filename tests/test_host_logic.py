"""CPU tests of the host-side mirror: constructor / state_dict / RNG-stream parity with the
reference, the C-ABI symbol table, and the no-CPU-fallback rule."""
import ctypes
import os
import re
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

from vqvae import _cabi, layers as L
from vqvae.model import VQVAE, full_config_args, downscaled_config_args


def test_full_config_shapes_and_param_count():
    torch.manual_seed(42)
    m = VQVAE(full_config_args())
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == 7_498_384                      # SURVEY.md section 6 [probe]
    blocks = [b for b in m.modules() if isinstance(b, L.PreActFixupResBlock)]
    assert len(blocks) == 361
    assert m.num_layers == 2 + 12 + 50 + 50 + 12 + 18 + 1
    q = m.encoder.quantize
    assert [(x.num_embeddings, x.embedding_dim) for x in q] == [(128, 2), (256, 8), (512, 32)]
    assert m.encoder.quantize[0].first_pass.dtype == torch.int64 and m.encoder.quantize[0].first_pass.dim() == 0


def test_downscaled_config_param_count():
    m = VQVAE(downscaled_config_args())
    assert sum(p.numel() for p in m.parameters()) == 965_856
    assert len([b for b in m.modules() if isinstance(b, L.PreActFixupResBlock)]) == 662


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("block_type", ["pre-activation", "regular"])
def test_same_seed_gives_reference_weights(block_type):
    """Constructors consume the RNG in the reference's order: seed 42 -> identical
    state_dict (names, order, values) as the reference's Encoder2/Decoder + Fixup init."""
    warnings.filterwarnings("ignore")
    import importlib.util
    # load the reference file under a private name (its `from vqvae.evonorm import ...` resolves to
    # this repo's evonorm module, which only matters for --block-type evonorm, not tested here)
    spec = importlib.util.spec_from_file_location("ref_layers_for_test", os.path.join(REF, "vqvae", "layers.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rb = {"pre-activation": ref.PreActFixupResBlock, "regular": ref.FixupResBlock}[block_type]
    torch.manual_seed(42)
    enc = ref.Encoder2(in_channels=1, base_network_channels=4, n_enc=2, n_down_per_enc=2, n_pre_q_blocks=2,
                       n_post_downscale_blocks=1, n_post_upscale_blocks=1, num_embeddings=[16, 32], resblock=rb)
    dec = ref.Decoder(out_channels=1, base_network_channels=4, n_enc=2, n_up_per_enc=2, n_post_q_blocks=2,
                      n_post_upscale_blocks=1, resblock=rb)
    holder = torch.nn.Module(); holder.encoder, holder.decoder = enc, dec
    holder.apply(lambda l: l.initialize_weights(num_layers=7) if isinstance(l, rb) else None)
    ref_sd = holder.state_dict()
    torch.manual_seed(42)
    args = VQVAE.default_args(n_bottleneck_blocks=2, num_embeddings=[16, 32], n_pre_quantization_blocks=2,
                              n_post_quantization_blocks=2, n_post_upscale_blocks=1, n_post_downscale_blocks=1,
                              block_type=block_type)
    m = VQVAE(args)
    torch.manual_seed(42)
    enc2 = L.Encoder2(in_channels=1, base_network_channels=4, n_enc=2, n_down_per_enc=2, n_pre_q_blocks=2,
                      n_post_downscale_blocks=1, n_post_upscale_blocks=1, num_embeddings=[16, 32], resblock=m.resblock)
    dec2 = L.Decoder(out_channels=1, base_network_channels=4, n_enc=2, n_up_per_enc=2, n_post_q_blocks=2,
                     n_post_upscale_blocks=1, resblock=m.resblock)
    h2 = torch.nn.Module(); h2.encoder, h2.decoder = enc2, dec2
    h2.apply(lambda l: l.initialize_weights(num_layers=7) if isinstance(l, m.resblock) else None)
    mine = h2.state_dict()
    assert list(mine.keys()) == list(ref_sd.keys())
    for k in ref_sd:
        assert mine[k].shape == ref_sd[k].shape and mine[k].dtype == ref_sd[k].dtype, k
        assert torch.equal(mine[k], ref_sd[k]), k


def test_cabi_header_and_binding_agree():
    """Every function declared in include/vqvae3d_b200.h is bound, and vice versa."""
    hdr = open(os.path.join(ROOT, "include", "vqvae3d_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vq3d_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)


def test_cabi_library_loads_and_exports_every_symbol():
    """The nvcc-built shared library loads without a GPU and exports the whole ABI
    (no compute calls here)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("vq3d_build", os.path.join(ROOT, "3d-vq-vae-2_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    so = b.build()
    lib = _cabi.declare(ctypes.CDLL(so))
    assert lib.vq3d_abi_version() == _cabi.ABI_VERSION
    assert lib.vq3d_is_cuda_build() == 1
    # argument validation runs before any CUDA call
    assert lib.vq3d_conv3d(None, None) == _cabi.ERR_INVALID
    assert b"null descriptor" in lib.vq3d_last_error()


def test_cabi_argument_validation_of_the_round2_entry_points():
    """Error behaviour of the new entry points (no GPU needed: validation runs before any CUDA call): null pointers, wrong
    block mode, geometry the fused kernels do not cover -> VQ3D_ERR_INVALID / VQ3D_ERR_UNSUPPORTED with a message, never a crash."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("vq3d_build", os.path.join(ROOT, "3d-vq-vae-2_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    lib = _cabi.declare(ctypes.CDLL(b.build()))
    C = ctypes
    assert lib.vq3d_preact_up_tc(None, None, 0, None) == _cabi.ERR_INVALID
    d = _cabi.PreactDesc(B=1, H=4, W=4, Z=4, Cin=18, Cb=9, Cout=8, mode=0)
    assert lib.vq3d_preact_up_tc(C.byref(d), None, 0, None) == _cabi.ERR_INVALID and b"mode must be 2" in lib.vq3d_last_error()
    assert lib.vq3d_preact_up_tc_workspace(C.byref(d)) == 0                       # not an 'up' block with a skip convolution
    d.mode = 2
    assert lib.vq3d_preact_up_tc(C.byref(d), None, 0, None) == _cabi.ERR_INVALID and b"null tensor" in lib.vq3d_last_error()
    assert lib.vq3d_conv1x1_backward(None, None, None) == _cabi.ERR_INVALID
    cd = _cabi.ConvDesc(B=1, H=4, W=4, Z=4, C1=4, C2=0, Cout=4, k=3, stride=1, pad=1, x1=1, w=1)
    g = _cabi.ConvBwd(gy=1)
    assert lib.vq3d_conv1x1_backward(C.byref(cd), C.byref(g), None) == _cabi.ERR_INVALID and b"pointwise" in lib.vq3d_last_error()
    s = _cabi.PreactDesc(B=1, H=4, W=4, Z=4, Cin=64, Cb=32, Cout=64, mode=0)
    assert lib.vq3d_preact_same_backward_workspace(C.byref(s)) == 0               # C > 32 / Cb > 16: composed path
    s = _cabi.PreactDesc(B=1, H=4, W=4, Z=4, Cin=18, Cb=9, Cout=18, mode=0)
    assert lib.vq3d_preact_same_backward_workspace(C.byref(s)) == 2 * 9 * 64 * 4
    assert lib.vq3d_preact_same_backward(C.byref(s), None, None, 0, None, None, None, None, None, None) == _cabi.ERR_INVALID
    assert lib.vq3d_hu_to_network(None, 4, -1500.0, 3000.0, 0.001, 1.0, None, None) == _cabi.ERR_INVALID
    assert lib.vq3d_elu_hu_rint_i16(None, 4, 1000.0, 1000.0, None, None) == _cabi.ERR_INVALID
    assert lib.vq3d_huber_elu_mask_stats(None, None, None, None, 1, 2, 2, 2, None, None, None) == _cabi.ERR_INVALID
    assert lib.vq3d_huber_elu_mask_medians_workspace() == 2 * 8 * 4 + 2 * 256 * 8
    assert lib.vq3d_huber_elu_mask_medians(None, None, None, None, 1, 2, 2, 2, None, None, 0, None) == _cabi.ERR_INVALID
    one = (C.c_float * 8)()
    assert lib.vq3d_huber_elu_mask_medians(one, one, None, None, 1, 2, 2, 2, one, one, 16, None) == _cabi.ERR_INVALID \
        and b"workspace" in lib.vq3d_last_error()


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing somewhere else."""
    q = L.Quantizer(8, 2, 0.1).eval()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        q(torch.zeros(1, 2, 2, 2, 2))
    blk = L.PreActFixupResBlock(4, 4, "same")
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA tensors only"):
        blk(torch.zeros(1, 4, 2, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):       # the training path refuses CPU tensors as well
        blk(torch.zeros(1, 4, 2, 2, 2))
    reg = L.FixupResBlock(4, 4, "same")
    with pytest.raises(RuntimeError, match="CUDA tensors only"):       # every block type trains through the CUDA kernels only
        reg(torch.zeros(1, 4, 2, 2, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "3d-vq-vae-2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "emu_ops" not in src, f


def test_scalar_gradient_scratch_pool_hands_out_zeroed_slots_and_falls_back():
    """Ops.begin_step / _gscal / end_step: slots of one pool zeroed by a single launch per step, distinct per call, re-zeroed by the
    next begin_step, and a private zeros(4) once the pool is spent or the step is over (a stale slot would corrupt gradients)."""
    from vqvae import _ops
    o = _ops.Ops.__new__(_ops.Ops)
    o._gscal_pool, o._gscal_next = {}, 0
    o.stream = lambda: 0
    dev = torch.device("cpu")
    a = o._gscal(dev)                                   # no begin_step yet: private scratch
    assert a.shape == (4,) and float(a.abs().sum()) == 0.0
    o.begin_step(dev, slots=2)
    s0, s1, s2 = o._gscal(dev), o._gscal(dev), o._gscal(dev)
    assert s0.data_ptr() != s1.data_ptr() and s0.untyped_storage().data_ptr() == s1.untyped_storage().data_ptr()
    assert s2.untyped_storage().data_ptr() != s0.untyped_storage().data_ptr()          # pool spent: fallback
    s0.add_(3.0); s1.add_(5.0)
    o.end_step()
    assert o._gscal(dev).untyped_storage().data_ptr() != s0.untyped_storage().data_ptr()
    o.begin_step(dev, slots=2)
    t0 = o._gscal(dev)
    assert t0.data_ptr() == s0.data_ptr() and float(t0.abs().sum()) == 0.0             # the same slot, zeroed again
