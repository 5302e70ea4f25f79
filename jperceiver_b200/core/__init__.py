"""Mirror of the reference's ``mono.core`` for the part SURVEY.md §8(f)-4 names: the validation hook and its metrics."""
from .evaluation import *  # noqa: F401,F403
