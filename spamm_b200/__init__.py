"""spamm_b200 -- B200-native ln-posterior hot path of SPAMM behind SPAMM's own
component plugin API.  See DESIGN.md."""
__version__ = "0.1.0"
