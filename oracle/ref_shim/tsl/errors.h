#pragma once
#include <tsl/result.h>
