"""gym_sbr2_b200 -- B200-native batched sequencing-batch-reactor simulator behind the gym-SBR API."""
__version__ = "0.1.0"
