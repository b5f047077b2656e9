"""Import-path shim: the reference's `project.*` module paths, served by recommendsystemproject_b200."""
