// MOCK (see conduit_node.hpp)
#pragma once
#include "conduit_blueprint.hpp"
