"""lycoris_b200 — B200-native adapter-layer engine behind the LyCORIS API."""
