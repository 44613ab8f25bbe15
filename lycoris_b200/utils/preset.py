"""User presets are toml files; an unreadable one is reported and ignored, as upstream does (lycoris/utils/preset.py:4-9)."""


def read_preset(preset):
    loader = _toml_loader()
    result = None
    if loader is not None:
        try:
            result = loader(preset)
        except Exception as err:  # noqa: BLE001 - bad path, bad syntax: same outcome upstream
            print("Error: cannot read preset file. ", err)
    return result


def _toml_loader():
    try:
        import toml
    except ImportError as err:
        print("Error: cannot read preset file. ", err)
        return None
    return toml.load
