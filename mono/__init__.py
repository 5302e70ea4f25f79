"""Import alias: ``mono.*`` resolves to ``jperceiver_b200.*`` so the reference's ``train.py`` imports
(``from mono.model.registry import MONO``, ``from mono.apis import train_mono, ...``) work unchanged."""
