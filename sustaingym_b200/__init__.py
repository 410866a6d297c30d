"""sustaingym_b200 — B200-native batched step() engine behind SustainGym's Gymnasium/PettingZoo API."""
__version__ = "0.1.0"
