"""Drop-in ``vision_base`` namespace: same dotted plugin names as the reference, B200-native implementations."""
