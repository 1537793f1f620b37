"""Drop-in ``monodepth`` namespace: same dotted plugin names as the reference, B200-native implementations."""
