#pragma once
// Forwarder: with rendu_b200/host/shadow ahead of src/engine on the include path, the engine's
// #include "raycaster/Raycaster.hpp" resolves to the GPU drop-in.
#include "../../Raycaster.hpp"
