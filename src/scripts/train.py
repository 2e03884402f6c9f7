"""`python -m src.scripts.train $(cat config/...)` — same entry point as the reference."""
from lipreading_b200.train_script import main, restore, train  # noqa: F401

if __name__ == "__main__":
    main()
