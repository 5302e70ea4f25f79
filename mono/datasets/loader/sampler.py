from jperceiver_b200.datasets.loader import DistributedGroupSampler, DistributedSampler, GroupSampler  # noqa: F401
