"""Import-path alias of the reference layout; the implementation lives in lipreading_b200/."""
