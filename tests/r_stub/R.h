#include <stddef.h>
