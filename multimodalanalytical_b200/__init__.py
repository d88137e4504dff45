"""B200-native (sm_100a) implementation of the spectra -> SMILES encoder-decoder hot path of
rxn4chemistry/MultimodalAnalytical (`analytical_fm`).  See DESIGN.md."""
__version__ = "0.1.0"
