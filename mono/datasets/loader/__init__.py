from jperceiver_b200.datasets.loader import (DistributedGroupSampler, DistributedSampler, GroupSampler, build_dataloader,  # noqa: F401
                                            collate)
