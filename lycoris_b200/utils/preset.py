def read_preset(preset):
    """Load a user preset from a toml file; None when unreadable (lycoris/utils/preset.py:4-9)."""
    try:
        import toml

        return toml.load(preset)
    except Exception as e:  # noqa: BLE001 - the reference reports and returns None
        print("Error: cannot read preset file. ", e)
        return None
