"""stub: the reference imports spacy at module load (data_loader.py:15); only --sentence_dataset uses it."""
