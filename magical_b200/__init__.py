"""magical_b200: B200-native batched implementation of the MAGICAL
(qxcv/magical) physics + render hot path.  See DESIGN.md."""
__version__ = '0.1.0'
