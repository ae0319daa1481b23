"""Word-language-model definition used by evaluate_lstm (lstm_models/model.py:6-62); the
reference's training script, generator and Transformer variant are out of scope."""
