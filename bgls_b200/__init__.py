"""bgls_b200: B200-native aggregate-signature verification engine (pairing products and
point aggregation for altbn128 / bls12-381) behind the Project-Arda/bgls curve interface."""
from ._native import ALTBN128, BLS12_381, G1, G2, BglsError, Context  # noqa: F401

__all__ = ["ALTBN128", "BLS12_381", "G1", "G2", "BglsError", "Context"]
