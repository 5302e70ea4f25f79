"""Block-level synchronisation of the CUDA-core kernels, checked without a GPU.

The default host emulation (``tests/emu``) runs one "thread" per block: it checks arithmetic and indexing but cannot see a
missing ``__syncthreads``, a wrong shuffle reduction or a shared-memory hand-off between threads.  Here the same kernel
sources are compiled with ``-DJPB_HOST_EMU_MT`` (``jperceiver_b200/csrc/jpb_common.cuh``): every block runs with its real
thread count on OS threads, barriers / shuffles / atomics behave as on the device, and a curated set of the existing parity
cases (small shapes, one per synchronisation pattern) is re-run against the same oracles.  The full emulation suites run
this way with ``JPB_EMU_MT=1 python -m pytest tests/<file> -m "not gpu"`` (minutes per file).

Patterns covered: block sums (shuffle + shared scratch, single and 24 in a row), staged shared-memory tiles read by other
threads, shared-memory atomics with a hand-off (disparity footprint, stream compaction), the radix-select broadcast loop,
last-block ticket counters (BatchNorm), warp-per-pixel shuffles (small-N convolution), early-exit threads.
TEST INFRASTRUCTURE ONLY: the product library is never built this way."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402

from jperceiver_b200 import _lib  # noqa: E402

import test_heads as TH  # noqa: E402
import test_losses as TL  # noqa: E402
import test_zz_eval as TE  # noqa: E402
import test_zz_pipeline as TP  # noqa: E402

CPU = torch.device("cpu")


@pytest.fixture(scope="module", autouse=True)
def threaded_emulation():
    _lib._handle, _lib._emulated = None, False
    _lib.use_library(build_emulation(mt=True), emulated=True)
    assert b"host-emulation" in _lib.lib().jpb_build_info()
    yield
    _lib.check(_lib.lib().jpb_photometric_set_variant(2), "jpb_photometric_set_variant")
    _lib._handle, _lib._emulated = None, False


@pytest.mark.parametrize("variant", [2, 3])
def test_photometric_forward_and_backward(variant):
    """Staged tiles + apron in shared memory, per-tile block sum, backward's shared-memory footprint atomics and 24 block sums."""
    _lib.check(_lib.lib().jpb_photometric_set_variant(variant), "jpb_photometric_set_variant")
    TL.test_photometric_kat5(CPU, variant)
    TL.test_photometric_forward_backward_vs_oracle(CPU, variant, 0, 17, 33, True)


def test_batchnorm_ticket_counter_and_block_sums():
    TL.test_batchnorm_train_forward_backward_vs_torch(CPU, 64, 12, 20, True, True)


def test_pool_smalln_and_loss_reductions():
    TL.test_maxpool_forward_backward_vs_torch(CPU, 5, 1, 2, 12, 20, 8)
    TL.test_small_n_conv_forward_backward_vs_torch(CPU, 256, 1, 1, 1, "sigmoid")      # one warp per pixel, xor-shuffle reduction
    TL.test_smoothness_forward_backward_vs_oracle(CPU, True)
    TL.test_l1_mean_vs_kat8(CPU)


def test_signed_distance_and_bev_loss():
    TL.test_signed_distance_exact(CPU)
    TL.test_bev_head_loss_forward_backward_vs_oracle(CPU, False)


def test_heads():
    TH.test_pose_head_forward_backward_vs_torch(CPU, False)
    TH.test_cct_attention_forward_backward_vs_torch(CPU)
    TH.test_cvp_mlp_forward_backward_vs_torch(CPU)


def test_eval_radix_select_and_compaction():
    """1024-thread block: histogram atomics, the scan-and-broadcast loop of the radix select, stream compaction through shared
    memory, seven block sums."""
    TE.test_depth_medians_are_exact_order_statistics(CPU)


def test_input_pipeline_reductions():
    TP.test_color_jitter_bit_exact_with_torchvision_on_pil(CPU, [2, 3, 0, 1], (1.2, 0.8, 1.2), 0.1)   # contrast mean: block sum + atomic
    TP.test_lanczos_resize_bit_exact_with_pillow(CPU, 33, 17, 11, 17)
