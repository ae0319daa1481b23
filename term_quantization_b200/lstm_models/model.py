import torch.nn as nn
import torch.nn.functional as F


class RNNModel(nn.Module):
    """Embedding -> dropout -> nn.LSTM/GRU/RNN -> dropout -> Linear decoder -> log-softmax.
    Attribute names (encoder, rnn, decoder, drop) follow lstm_models/model.py:6-62 because
    evaluate_lstm replaces `rnn` and `decoder` by name order."""

    def __init__(self, rnn_type, ntoken, ninp, nhid, nlayers, dropout=0.5, tie_weights=False):
        super().__init__()
        self.ntoken = ntoken
        self.drop = nn.Dropout(dropout)
        self.encoder = nn.Embedding(ntoken, ninp)
        if rnn_type in ('LSTM', 'GRU'):
            self.rnn = getattr(nn, rnn_type)(ninp, nhid, nlayers, dropout=dropout)
        elif rnn_type in ('RNN_TANH', 'RNN_RELU'):
            self.rnn = nn.RNN(ninp, nhid, nlayers, dropout=dropout,
                              nonlinearity=rnn_type.split('_')[1].lower())
        else:
            raise ValueError("--model must be one of LSTM, GRU, RNN_TANH, RNN_RELU")
        self.decoder = nn.Linear(nhid, ntoken)
        if tie_weights:
            if nhid != ninp:
                raise ValueError('When using the tied flag, nhid must be equal to emsize')
            self.decoder.weight = self.encoder.weight
        self.rnn_type, self.nhid, self.nlayers = rnn_type, nhid, nlayers
        self.init_weights()

    def init_weights(self):
        self.encoder.weight.data.uniform_(-0.1, 0.1)
        self.decoder.bias.data.zero_()
        self.decoder.weight.data.uniform_(-0.1, 0.1)

    def forward(self, input, hidden):
        emb = self.drop(self.encoder(input))
        output, hidden = self.rnn(emb, hidden)
        decoded = self.decoder(self.drop(output)).view(-1, self.ntoken)
        return F.log_softmax(decoded, dim=1), hidden

    def init_hidden(self, bsz):
        w = next(self.parameters())
        zeros = w.new_zeros(self.nlayers, bsz, self.nhid)
        return (zeros, zeros.clone()) if self.rnn_type == 'LSTM' else zeros
