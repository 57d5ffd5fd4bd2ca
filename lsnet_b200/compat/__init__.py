"""Drop-in shims with the reference's own native-extension interfaces."""
