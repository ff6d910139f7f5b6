"""Registry of the accelerated confs -- same shape as the reference's imcui/hloc/configs/__init__.py:1-7:
confs_dict = {"extractors": {...}, "matchers": {...}}, each conf = {"output", "model": {"name", ...},
"preprocessing": {...}}.  Only the north-star confs are listed; values are the reference's."""
from .extractors import confs as extractors_confs
from .matchers import confs as matchers_confs

confs_dict = {
    "extractors": extractors_confs,
    "matchers": matchers_confs,
}
