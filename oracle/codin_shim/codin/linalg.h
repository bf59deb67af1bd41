/* TEST INFRASTRUCTURE: part of the Codin stand-in, see codin.h */
#include "codin.h"
