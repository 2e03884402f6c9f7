"""README.md:180 of the reference still names `train_model.py` (archived there); alias of train.py."""
from lipreading_b200.train_script import main, train  # noqa: F401

if __name__ == "__main__":
    main()
