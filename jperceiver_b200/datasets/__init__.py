"""Device-side input pipeline: the per-sample PIL work of ``mono/datasets/mono_dataset.py`` as batched kernels."""
from .preprocess import (GpuPreprocess, bev_label, color_jitter, draw_color_jitter, lanczos_tables, nearest_table,  # noqa: F401
                         resize_lanczos)
from .loader import (DistributedGroupSampler, DistributedSampler, GroupSampler, build_dataloader, collate,  # noqa: F401
                     get_dist_info)
