"""mr_blip_b200: B200-native (sm_100a) implementation of the BLIP2_MR / Chrono per-step hot path
behind the reference's LAVIS model surface.  See DESIGN.md."""
__version__ = "0.1.0"
