/* forwarder: a model file written for APEMoST includes "markov_chain.h"; everything it can use from it is in apm_host.h */
#include "apm_host.h"
