"""afcm_b200 -- B200-native (sm_100a) implementation of the AFCM alias-free co-modulated generator
forward path behind the reference's operator API.  See DESIGN.md."""
__version__ = '0.1.0'
