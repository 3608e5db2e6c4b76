"""gbrs_b200 -- B200-native multiway EM quantifier (the loop behind `gbrs quantify -M 1..4`).

Public surface (mirrors churchill-lab/gbrs for this path):
    AlignmentPropertyMatrix   container for a compressed-EMASE incidence matrix
    EMfactory                 prepare / run / report on the GPU
    quantify                  the `gbrs quantify` workflow
"""
from .apm import AlignmentPropertyMatrix  # noqa: F401
from .emfactory import EMfactory  # noqa: F401
from .quantify import quantify  # noqa: F401

__version__ = "0.1.0"
