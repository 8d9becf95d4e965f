"""Drivers with the reference's surface: ``maml_vision.MamlVision`` and ``anil_vision.AnilVision``
(reference ``vision/maml_vision.py`` / ``vision/anil_vision.py``)."""
