"""starcop_b200 -- B200-native (sm_100a) implementation of STARCOP's data-parallel hot path."""
