from jperceiver_b200.apis import get_root_logger, init_dist, set_random_seed, train_mono  # noqa: F401
